"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads without a GPU driver,
exports every symbol include/uzu_b200.h declares, mirrors the struct layouts, and fails loudly without a GPU."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(lib):
    from uzu_b200 import binding
    header = (ROOT / "include" / "uzu_b200.h").read_text()
    declared = set(re.findall(r"UZU_API\s+[\w\s\*]+?\b(uzu_\w+)\s*\(", header))
    assert len(declared) > 60
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"library does not export: {missing}"
    assert declared == set(binding.EXPORTS), (declared ^ set(binding.EXPORTS))


def test_struct_layouts_match(lib):
    import ctypes as C
    from uzu_b200 import binding
    for st in binding.ABI_STRUCTS:
        assert lib.uzu_abi_struct_size(st.__name__.encode()) == C.sizeof(st), st.__name__
    assert lib.uzu_abi_struct_size(b"no_such_struct") == 0


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to run (it never routes to oracle/ or any CPU path)."""
    import ctypes as C
    h = C.c_void_p()
    st = lib.uzu_context_create(0, C.byref(h))
    if st == 0:   # a GPU is present (GPU box): nothing to assert here
        lib.uzu_context_destroy(h)
        pytest.skip("GPU present")
    assert st == 6 and b"no CPU fallback" in lib.uzu_last_error()


def test_product_does_not_import_oracle():
    for py in (ROOT / "uzu_b200").glob("*.py"):
        text = py.read_text()
        assert "import oracle" not in text and "from oracle" not in text, py
    for src in (ROOT / "uzu_b200" / "csrc").glob("*"):
        assert "liboracle" not in src.read_text() and "uzu_oracle" not in src.read_text(), src


def test_sass_uses_tensor_pipe_and_vector_loads(lib):
    """The decode GEMV is built around mma.sync (HMMA in SASS) and 128-bit streaming loads."""
    import shutil, subprocess
    from uzu_b200 import binding
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(binding.LIB_PATH)], capture_output=True, text=True).stdout
    assert "HMMA.16816.F32.BF16" in sass
    assert "LDG.E.128" in sass


def test_sass_of_the_prefill_gemm_uses_tcgen05(lib):
    """The prefill GEMM runs on the 5th-generation tensor cores: tcgen05.mma (UTCHMMA), TMEM loads (LDTM), tcgen05.commit (UTCBAR),
    TMEM allocation (UTCATOMSWS) and cp.async staging (LDGSTS) must be in the sm_100a SASS (B200_PROFILING.md mnemonics)."""
    import shutil, subprocess
    from uzu_b200 import binding
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3uzu15qmm_umma_kernelILi4ELi2ELi3ELi256ELb1EEEvNS_9QmmParamsE", str(binding.LIB_PATH)],
                          capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "LDGSTS", "HFMA2.BF16"):
        assert mnemonic in sass, mnemonic


def test_header_is_plain_c(tmp_path):
    """include/uzu_b200.h is the drop-in boundary: it must compile as C11 for cgo / bindgen / ctypes-style consumers."""
    import shutil, subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "h.c"
    src.write_text('#include "uzu_b200.h"\nint main(void) { return sizeof(uzu_matmul_args) + sizeof(uzu_tp_all_gather_args) > 0 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "h.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
