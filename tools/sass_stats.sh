#!/bin/bash
# static SASS instruction count per kernel of libbetse_b200.so
d=$(mktemp -d); cd $d; cuobjdump -xelf all /root/repo/betse_b200/libbetse_b200.so >/dev/null
for f in *.cubin; do nvdisasm -c $f; done | python3 -c "
import sys,re,collections
cur=None;cnt=collections.Counter()
for ln in sys.stdin:
    m=re.match(r'\t\.section\t\.text\.(\S+?),',ln)
    if m: cur=m.group(1); continue
    if ln.startswith('\t.section'): cur=None
    if cur and re.match(r'\s*/\*[0-9a-f]{4,}\*/',ln): cnt[cur]+=1
for k,v in sorted(cnt.items()): print(v,k)
"
rm -rf $d
