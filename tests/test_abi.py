"""The C-ABI library loads and exports every symbol include/insb200.h declares (no GPU needed)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "insb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(insb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import insb200

    assert os.path.exists(insb200.LIB_PATH), "build the library first (__graft_entry__.build())"
    L = ctypes.CDLL(insb200.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in insb200.h but not exported"
    assert sorted(insb200.EXPORTED_SYMBOLS) == names
    assert insb200.lib().insb200_version() == 100


def test_setup_struct_matches_header_layout():
    # sizeof(insb200_setup_t) as the C compiler sees it == ctypes layout
    import subprocess, tempfile

    from insb200._lib import ParamsStruct, SetupStruct

    code = '#include <stdio.h>\n#include "insb200.h"\nint main(){printf("%zu %zu\\n", sizeof(insb200_setup_t), sizeof(insb200_params_t));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "a.c"), "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "a.c"), "-o", os.path.join(d, "a")])
        out = subprocess.check_output([os.path.join(d, "a")]).split()
    assert int(out[0]) == ctypes.sizeof(SetupStruct)
    assert int(out[1]) == ctypes.sizeof(ParamsStruct)


def test_plan_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        return
    import insb200
    from insb200._lib import SetupStruct

    s = SetupStruct()
    s.D = 2
    h = ctypes.c_void_p()
    rc = insb200.lib().insb200_plan_create(ctypes.byref(s), ctypes.byref(h))
    assert rc != 0
    assert b"no CUDA device" in insb200.lib().insb200_last_error()


def test_host_setup_matches_oracle_tables():
    # host logic: the product's Setup builds the same tables as the oracle's restatement
    import insb200
    import oracle as O

    x = (O.tanh_grid(0, 1, 12, 1.2), np.linspace(0, 2, 9), O.cosine_grid(0, 1, 10))
    combos = [
        ((insb200.DirichletBC(), insb200.DirichletBC()), (O.DirichletBC(), O.DirichletBC())),
        ((insb200.PeriodicBC(), insb200.PeriodicBC()), (O.PeriodicBC(), O.PeriodicBC())),
        ((insb200.SymmetricBC(), insb200.PressureBC()), (O.SymmetricBC(), O.PressureBC())),
    ]
    for T in (np.float32, np.float64):
        xs = tuple(v.astype(T) for v in x)
        a = insb200.Setup(xs, {"u": tuple(c[0] for c in combos)})
        b = O.Setup(xs, {"u": tuple(c[1] for c in combos)})
        assert a.N == b.N and a.Np == b.Np
        for d in range(3):
            assert a.Ip[d] == (b.Ip[d].start, b.Ip[d].stop)
            for c in range(3):
                assert a.Iu[c][d] == (b.Iu[c][d].start, b.Iu[c][d].stop)
            for name in ("Delta", "DeltaU", "A1", "A2", "x", "xp"):
                assert np.array_equal(getattr(a, name)[d], getattr(b, name)[d]), name
