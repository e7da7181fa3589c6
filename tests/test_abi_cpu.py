"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol the header declares, and the compute
entry points refuse to run without an sm_100 device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "abcnet_b200.h")).read()
    return sorted(set(re.findall(r"ABC_API[^;(]*?\b(abc_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    from abcnet_b200 import _lib
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/abcnet_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names)
    assert _lib.lib.abc_version() >= 100


def test_struct_layouts_match_header(tmp_path):
    """Every ctypes mirror in abcnet_b200/_lib.py has the size and field offsets the C compiler gives the struct declared in
    include/abcnet_b200.h (gcc compiles a probe that prints sizeof / offsetof)."""
    import subprocess
    from abcnet_b200 import _lib
    assert C.sizeof(_lib.AbcAtomRec) == 8 and C.sizeof(_lib.AbcBondRec) == 12
    structs = [n for n in dir(_lib) if n.startswith("Abc") and isinstance(getattr(_lib, n), type) and issubclass(getattr(_lib, n), C.Structure)]
    assert {"AbcConvDesc", "AbcDecodeDesc", "AbcLossDesc", "AbcHeadsFusedDesc", "AbcWgradDesc", "AbcBnActDesc", "AbcBnActBwdDesc"} <= set(structs)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "abcnet_b200.h"', 'int main(void) {']
    for sname in structs:
        lines.append(f'  printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, *_ in getattr(_lib, sname)._fields_:
            lines.append(f'  printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname.rstrip("_")}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for sname in structs:
        cls = getattr(_lib, sname)
        assert int(got[sname]) == C.sizeof(cls), f"sizeof({sname}): C {got[sname]} vs ctypes {C.sizeof(cls)}"
        for fname, *_ in cls._fields_:
            assert int(got[f"{sname}.{fname}"]) == getattr(cls, fname).offset, f"offsetof({sname}, {fname})"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    import abcnet_b200
    from abcnet_b200 import _lib
    assert _lib.lib.abc_device_ok() == 0
    d = _lib.AbcDecodeDesc()
    assert _lib.lib.abc_decode_peaks(C.byref(d), None) == -2            # ABC_ERR_NO_DEVICE
    assert b"no CPU fallback" in _lib.lib.abc_last_error() or b"CUDA" in _lib.lib.abc_last_error()
    m = abcnet_b200.UNet(1, [1, 14, 3, 2, 1, 360, 60, 60]).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 32, 32))
    m3 = abcnet_b200.UNet(3, [1, 14]).eval()                               # unet.py:127: the reference's self-check model
    assert m3.inc1.double_conv[0].weight.shape == (16, 3, 3, 3)
    with pytest.raises(RuntimeError):
        m3(torch.zeros(1, 3, 32, 32))
    # product code never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "abcnet_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
    # ... nor do the tools; bench.py only inside its CPU-baseline / reference-arm functions
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle import", bench)]
    gpu_arm = bench.index("def run_train(")
    assert uses and all(u < gpu_arm for u in uses), "bench.py: the oracle may only appear in cpu_path / calibrate_offsets_cpu"


def test_state_dict_matches_reference_inventory():
    import abcnet_b200
    from oracle import unet_ref
    m = abcnet_b200.UNet(1, list(unet_ref.V2_HEADS))
    ref = unet_ref.param_shapes()
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert list(mine.keys()) == list(ref.keys()) and mine == dict(ref)
