"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/betse_b200.h declares, the ctypes mirrors agree with the header's structs, and the
product path fails loudly without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "betse_b200.h")


def _header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(betse_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from betse_b200 import capi
    lib = capi.load()
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libbetse_b200.so does not export %s" % s
    assert sorted(capi.SYMBOLS) == syms, (sorted(capi.SYMBOLS), syms)
    assert lib.betse_abi_version() == capi.ABI_VERSION


def test_ctypes_structs_match_header_layout():
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    from betse_b200 import capi
    probes = [("betse_mesh", "memsa_mean", capi.Mesh), ("betse_mesh", "ecm_slot_idx", capi.Mesh),
              ("betse_params", "T_sim", capi.Params), ("betse_params", "gauss_w", capi.Params),
              ("betse_state_host", "cenv_uniform", capi.StateHost),
              ("betse_channel", "time_unit", capi.Channel), ("betse_channel", "h0", capi.Channel),
              ("betse_channel", "mod_prog", capi.Channel),
              ("betse_network", "c_cells", capi.Network), ("betse_network", "time_factor", capi.Network),
              ("betse_neighbor", "v_dst_row0", capi.Neighbor), ("betse_window_info", "off_flags", capi.WindowInfo)]
    body = "".join('printf("%%zu %%zu\\n", sizeof(%s), offsetof(%s, %s));\n' % (s, s, f) for s, f, _ in probes)
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "betse_b200.h"\nint main(){%s return 0;}\n' % body
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "p.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "p")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    for (s, f, cls), line in zip(probes, out):
        size, off = (int(x) for x in line.split())
        assert C.sizeof(cls) == size, (s, C.sizeof(cls), size)
        assert getattr(cls, f).offset == off, (s, f, getattr(cls, f).offset, off)


def test_header_compiles_as_plain_c():
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "h.c")
        open(src, "w").write('#include "betse_b200.h"\nint main(void){return BETSE_ABI_VERSION==0;}\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-c", src, "-o", os.path.join(td, "h.o")], check=True)


def test_no_cpu_fallback():
    """Without a CUDA device the engine must refuse to construct (it never computes on the host)."""
    from betse_b200 import capi, BetseB200Error
    lib = capi.load()
    if lib.betse_device_count() > 0:
        pytest.skip("a CUDA device is present")
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    mesh, p, st = synth.make_tissue(200)
    with pytest.raises(BetseB200Error):
        TissueEngine(mesh, p, st)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under betse_b200/ may import it."""
    pkg = os.path.join(ROOT, "betse_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), fn


def test_host_helpers_without_a_gpu():
    """betse_host_copy / betse_host_expand are plain host code (threaded memcpy / row gather used by the copy-back of
    Simulator arrays): exact against NumPy, odd sizes and sizes below and above the threading thresholds."""
    from betse_b200 import capi
    lib = capi.load()
    rng = np.random.default_rng(0)
    for n in (0, 1, 1000, (1 << 18) + 3, (3 << 20) + 17):
        src = rng.random(n)
        dst = np.full(n, -1.0)
        lib.betse_host_copy(dst.ctypes.data, src.ctypes.data, C.c_size_t(src.nbytes))
        assert np.array_equal(dst, src)
    for n_src, rows in ((5, 1), (1000, 4), (70_000, 6)):
        counts = rng.integers(3, 8, n_src)
        idx = np.repeat(np.arange(n_src), counts).astype(np.int32)
        src = rng.random((rows, n_src))
        dst = np.empty((rows, len(idx)))
        lib.betse_host_expand(capi.ptr_f64(dst), capi.ptr_f64(src), capi.ptr_i32(idx), rows, n_src, len(idx))
        assert np.array_equal(dst, src[:, idx])


def test_staging_prefetch_degrades_without_a_gpu():
    """The drop-in loop pins its sample staging on a helper thread (simloop._prefetch_staging -> engine.PinnedPrefetch):
    the field list and shapes are those of the per-sample download, and without a CUDA device the thread simply hands
    back nothing (the engine then allocates — and reports — on first use)."""
    import types
    from betse_b200 import simloop, synth
    mesh, p, state = synth.make_tissue(500)
    cells = types.SimpleNamespace(cell_vol=mesh["cell_vol"], mem_sa=mesh["mem_sa"], X=None,
                                  grid_shape=tuple(int(x) for x in mesh["grid_shape"]))
    pp = types.SimpleNamespace(**p)
    pf = simloop._prefetch_staging(cells, pp, 0)
    assert pf is not None
    I, Cn, M, E = len(p["ions"]), len(mesh["cell_vol"]), len(mesh["mem_sa"]), int(np.prod(mesh["grid_shape"]))
    assert set(pf.shapes) == set(simloop._sample_fields(True, True))
    assert pf.shapes["cc_cells"] == (I, Cn) and pf.shapes["vm"] == (M,) and pf.shapes["cc_env"] == (I, E)
    assert pf.shapes["J_env_x"] == (E,) and pf.shapes["I_mem"] == (M,)
    got = pf.join()
    assert got == {} or set(got) <= set(pf.shapes)           # {} here; on a GPU box the arrays
    pf.release()
    pp.is_ecm = False
    assert simloop._prefetch_staging(cells, pp, 0) is None     # tissues without extracellular spaces: allocated on first use
