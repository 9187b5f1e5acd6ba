"""The round-2 micro-benchmark (tests/probe_cta2.cu) must cross-compile for sm_100a and really contain
CTA-pair tensor-core instructions (no GPU needed: nvcc + cuobjdump only)."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(shutil.which("nvcc") is None and not Path("/usr/local/cuda/bin/nvcc").exists(), reason="no nvcc")
def test_probe_cta2_compiles_to_2cta_mma():
    subprocess.run(["bash", str(ROOT / "tools" / "build_probe.sh")], check=True, capture_output=True)
    exe = ROOT / "tests" / "_probe" / "probe_cta2"
    assert exe.exists()
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([cuobjdump, "-sass", str(exe)], check=True, capture_output=True, text=True).stdout
    assert "UTCHMMA.2CTA" in sass          # tcgen05.mma.cta_group::2
    assert "UTCBAR.2CTA.MULTICAST" in sass  # tcgen05.commit ... multicast::cluster
    assert "UTCHMMA " in sass or "UTCHMMA\n" in sass or sass.count("UTCHMMA") > sass.count("UTCHMMA.2CTA")  # single-CTA lines too
    # the protocol probe: 2-CTA TMA loads, pair MMA, multicast commits (tests/probe_cta2_tma.cu)
    exe2 = ROOT / "tests" / "_probe" / "probe_cta2_tma"
    assert exe2.exists()
    sass2 = subprocess.run([cuobjdump, "-sass", str(exe2)], check=True, capture_output=True, text=True).stdout
    for mnemonic in ("UTMALDG.2D.2CTA", "UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in sass2, mnemonic
    # the stand-alone CTA-pair scoring sweep (tests/probe_cta2_sweep.cu)
    exe3 = ROOT / "tests" / "_probe" / "probe_cta2_sweep"
    assert exe3.exists()
    sass3 = subprocess.run([cuobjdump, "-sass", str(exe3)], check=True, capture_output=True, text=True).stdout
    for mnemonic in ("UTMALDG.2D.2CTA", "UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST"):
        assert mnemonic in sass3, mnemonic
