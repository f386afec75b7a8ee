"""The C-ABI library loads without a GPU and exports every symbol include/msd.h declares (no compute calls here)."""
import ctypes
import re
import subprocess
from pathlib import Path

import msd_b200
from msd_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "msd.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 20
    lib = _lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in msd.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_exports_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared_symbols()) <= exported
    # no libcuda / libtorch link dependency: the library must load on a driverless machine
    needed = subprocess.run(["readelf", "-d", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libcuda.so" not in needed and "libtorch" not in needed and "libc10" not in needed


def test_version_and_host_only_entry_points():
    lib = _lib.load()
    assert lib.msd_version() == 2
    # ((numel-1)/(256*grid*4)+1)*4  (ATen DistributionTemplates.h:60)
    assert lib.msd_philox_offset_increment(13, 1) == 4
    assert lib.msd_philox_offset_increment(256 * 13 * 192 * 192, 1184) == 408
    assert lib.msd_philox_offset_increment(1, 1) == 4
    assert lib.msd_loss_workspace_bytes(2, 13, 192, 192) == 256 + 3 * 2 * 13 * 192 * 192 * 4


def test_invalid_arguments_fail_loudly_without_touching_a_device():
    lib = _lib.load()
    rc = lib.msd_metrics_f32(None, None, 1, 13, 8, 8, 6.0, None, None, None, None)
    assert rc == -1 and b"null" in lib.msd_last_error()
    rc = lib.msd_attention_bf16(ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 100, 12, None)
    assert rc == -1 and b"T=144" in lib.msd_last_error()
    h = ctypes.c_void_p()
    d = _lib.ModelDesc()
    d.embed_dim, d.num_heads = 768, 10
    assert lib.msd_model_create(ctypes.byref(d), ctypes.byref(h)) == -1


def test_product_modules_never_import_the_oracle():
    pkg = ROOT / "multi-spectral-satellite-image-denoising_b200"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
