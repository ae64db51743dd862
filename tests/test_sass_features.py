"""The built library is sm_100a code that uses the Blackwell data path -- checked on the build host from the SASS of the
in-tree libsafconv_b200.so (cuobjdump, no GPU): TMA bulk copies + mbarriers in the MAC, tcgen05 MMA / TMEM loads in the
offline GEMM, cluster barriers and distributed-shared-memory stores in the cluster kernels, fp64 reductions in the
image-source render.  (profiles/rNN_sass_summary.txt is the per-kernel table of the same dump, tools/sass_summary.py.)"""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "spatial_audio_framework_b200" / "libsafconv_b200.so"


@pytest.fixture(scope="module")
def sass():
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cu).exists() or not LIB.exists():
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run([cu, "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); per[cur] = []
        elif cur is not None:
            per[cur].append(line)
    assert "sm_100a" in out or "EF_CUDA_SM100" in out or "arch = sm_100" in out
    return {k: "\n".join(v) for k, v in per.items()}


def kernels(sass, name):
    """bodies of the kernels whose (mangled) function name is exactly `name`, all template instances"""
    ks = [v for k, v in sass.items() if re.match(rf"_Z\d+{name}(I|v|P|\d|[A-Z])", k) and re.match(rf"_Z{len(name)}{name}", k)]
    assert ks, f"no kernel named {name} in the library"
    return ks


def test_mac_kernel_streams_with_tma_and_mbarriers(sass):
    for body in kernels(sass, "mac_kernel"):
        assert "UBLKCP" in body and "SYNCS" in body and "FFMA" in body


def test_offline_gemm_uses_tcgen05_and_tmem(sass):
    for body in kernels(sass, "offline_gemm_kernel"):
        assert "UTCHMMA" in body and "LDTM" in body and "UTCBAR" in body and "UBLKCP" in body


def test_cluster_kernels_use_cluster_barriers_and_dsmem(sass):
    for name in ("small_cluster_kernel", "small_cluster_resident_kernel", "prod_magls_cluster_kernel"):
        for body in kernels(sass, name):
            assert "UCGABAR_ARV" in body and "UCGABAR_WAIT" in body, name       # barrier.cluster.arrive / .wait
    # the MagLS kernel hands its partial decoders over with 16-byte stores into the other CTAs' shared memory
    # (st.shared::cluster.v2.f64 = a generic ST.E.128 into the cluster window; its global stores are 8-byte float2)
    for body in kernels(sass, "prod_magls_cluster_kernel"):
        assert "ST.E.128" in body


def test_image_source_render_reduces_in_fp64(sass):
    for body in kernels(sass, "ims_render_kernel"):
        assert re.search(r"RED\S*\.ADD\S*\.F64|RED\.E\.ADD\.F64", body)
    for body in kernels(sass, "ims_window_kernel"):
        assert "ATOMS" in body                                          # shared-memory accumulation, no RED to HBM
        assert not re.search(r"RED\S*\.ADD\S*\.F64", body)


def test_no_library_kernels_inside(sass):
    """everything in the .so is this repo's: no cuFFT / cuBLAS / CUTLASS device code linked in"""
    for k in sass:
        assert not re.search(r"cufft|cublas|cutlass|cudnn", k, re.I), k
