for v in "BETSE_CREATE_THREADS=1 BETSE_PIN_PREFETCH=1" "BETSE_CREATE_THREADS=0 BETSE_PIN_PREFETCH=1" "BETSE_CREATE_THREADS=1 BETSE_PIN_PREFETCH=0" "BETSE_CREATE_THREADS=0 BETSE_PIN_PREFETCH=0"; do
  for rep in 1 2 3; do echo "== $v"; env $v python tools/e2e_cprofile.py 20 2>&1 | grep -E "^rep" ; done
done
