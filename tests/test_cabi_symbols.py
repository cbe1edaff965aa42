"""CPU: the C-ABI library builds, loads and exports every symbol include/focal_b200.h declares (no GPU calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from focal_b200 import _cabi, build
    build.build()                                     # no-op when the in-tree .so is up to date
    return _cabi.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "focal_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(focal_b200_\w+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ("focal_b200_workspace_info", "focal_b200_prologue", "focal_b200_nce_rowsum", "focal_b200_nce_lse",
                 "focal_b200_nce_grad", "focal_b200_temporal", "focal_b200_finalize", "focal_b200_loss",
                 "focal_b200_strerror", "focal_b200_abi_version"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from focal_b200 import _cabi
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/focal_b200.h but not exported"
    assert set(_cabi.EXPORTS) <= set(declared_functions())
    assert lib.focal_b200_abi_version() == _cabi.ABI_VERSION


def test_struct_layouts_match_the_header():
    from focal_b200 import _cabi
    # FocalCfg: 4 + 6 + 7 + 4 32-bit fields
    assert C.sizeof(_cabi.FocalCfg) == 4 * 21
    text = open(os.path.join(ROOT, "include", "focal_b200.h")).read()
    body = re.search(r"typedef struct FocalCfg \{(.*?)\} FocalCfg;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(None, 1)[1]
        for n in names.split(","):
            fields.append(re.sub(r"\[.*\]", "", n.strip()))
    assert fields == [f[0] for f in _cabi.FocalCfg._fields_]


def test_ctypes_structs_match_the_compiled_header(tmp_path):
    """Sizes and field offsets of every struct of include/focal_b200.h as a C compiler lays them out == the ctypes mirror."""
    import shutil
    import subprocess
    from focal_b200 import _cabi
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {"FocalCfg": _cabi.FocalCfg, "FocalWsInfo": _cabi.FocalWsInfo, "FocalPeers": _cabi.FocalPeers}
    lines = []
    for sname, cls in structs.items():
        lines.append(f'printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname}));')
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "focal_b200.h"\nint main(void) {\n'
                   + "\n".join(lines) + "\nreturn 0; }\n")
    exe = str(tmp_path / "layout")
    res = subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True).stdout.splitlines())
    for sname, cls in structs.items():
        assert int(out[sname]) == C.sizeof(cls), sname
        for fname, _ in cls._fields_:
            assert int(out[f"{sname}.{fname}"]) == getattr(cls, fname).offset, f"{sname}.{fname}"


def test_workspace_info_and_error_codes_without_a_gpu(lib):
    from focal_b200 import _cabi

    def cfg(**kw):
        base = dict(B=8192, S=4, M=2, D=256, temperature=0.5, margin=1.0, w_shared=1.0, w_private=1.0, w_orth=3.0,
                    w_rank=5.0, no_private=0, need_grad=1, terms=7, precision=0, seq_begin=0, seq_end=2048, num_sms=148)
        base.update(kw)
        return _cabi.FocalCfg(**base)

    info = _cabi.FocalWsInfo()
    assert lib.focal_b200_workspace_info(C.byref(cfg()), C.byref(info)) == 0
    assert info.b == 2048 and info.n_problems == 4 and info.n_ops == 8 and info.kb_full == 4
    assert info.bpad % 128 == 0 and info.bpad >= 2048 and info.Bpad % 128 == 0 and info.Bpad >= 8192
    assert info.total_bytes % 1024 == 0 and 50e6 < info.total_bytes < 200e6
    assert info.rowsum_bytes == 4 * 4 * 2 * info.bpad * 4
    # M^2 problems, 4M operands
    c3 = cfg(M=3, B=4096, seq_end=1024)
    assert lib.focal_b200_workspace_info(C.byref(c3), C.byref(info)) == 0 and info.n_problems == 9 and info.n_ops == 12
    # shape errors the Python layer turns into ValueError
    for bad in (cfg(B=8190), cfg(S=33, B=8184, seq_end=248), cfg(D=514), cfg(M=9), cfg(temperature=0.01),
                cfg(precision=1, D=320)):
        assert lib.focal_b200_workspace_info(C.byref(bad), C.byref(info)) == _cabi.FOCAL_ESHAPE
    # shapes the reference accepts and round 1 refused: any seq_len up to 32 (padded to a power of two inside), M up to 8
    assert lib.focal_b200_workspace_info(C.byref(cfg(S=3, B=8190, seq_end=2730)), C.byref(info)) == 0
    assert info.Bpad >= 2730 * 4                                    # temporal row space: sequences padded to 4 rows
    assert lib.focal_b200_workspace_info(C.byref(cfg(M=5)), C.byref(info)) == 0 and info.n_problems == 25
    for bad in (cfg(seq_begin=5, seq_end=5), cfg(seq_end=4096), cfg(temperature=0.0)):
        assert lib.focal_b200_workspace_info(C.byref(bad), C.byref(info)) == _cabi.FOCAL_EINVAL
    assert lib.focal_b200_workspace_info(None, C.byref(info)) == _cabi.FOCAL_EINVAL
    assert b"unsupported shape" in lib.focal_b200_strerror(_cabi.FOCAL_ESHAPE)
    assert lib.focal_b200_strerror(0) == b"ok"


def test_missing_library_fails_loudly(tmp_path):
    from focal_b200 import _cabi
    with pytest.raises(ImportError):
        _cabi.load(str(tmp_path / "libfocal_b200.so"))


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The shipped binary is the tcgen05/TMEM/TMA path, not a legacy mma.sync recompile."""
    import shutil
    import subprocess
    from focal_b200 import _cabi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UBLKCP" in sass and "LDTM" in sass and "STTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")
