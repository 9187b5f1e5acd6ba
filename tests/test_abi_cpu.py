"""CPU-side checks: the C-ABI library builds/loads, exports every symbol include/recboard_b200.h
declares, and the product path refuses to run without a GPU (no silent fallback)."""
import re
from pathlib import Path

import pytest
import torch

from recboard_b200 import _lib as L
from recboard_b200 import build as rb_build

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    rb_build.build()
    return L.lib()


def test_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "recboard_b200.h").read_text()
    declared = set(re.findall(r"\b(rb_[a-z_0-9]+)\s*\(", header))
    declared.discard("rb_stream_t")
    assert declared, "no declarations parsed"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None


def test_version_and_workspace(lib):
    assert b"sm_100a" in lib.rb_version()
    for op in (L.OP_SCORE_DENSE, L.OP_CE_FWD, L.OP_CE_BWD, L.OP_TOPK_EVAL):
        assert L.workspace_bytes(op, 4096, 1_000_000, 128, K=50) > 0
    assert L.workspace_bytes(L.OP_SCATTER_ADD, 0, 0, 128, nnz=204800) >= 204800 * 16
    # fp32x3 stages [hi|lo] copies of both operands
    assert L.workspace_bytes(L.OP_CE_FWD, 512, 12101, 64, mode=L.MODE_FP32X3) > 12101 * 64 * 8


def test_no_cpu_fallback():
    from recboard_b200 import ops
    U, W = torch.zeros(8, 64), torch.zeros(16, 64)
    lab = torch.zeros(8, dtype=torch.int64)
    idx = torch.zeros(8, 3, dtype=torch.int64)
    A = torch.eye(16).to_sparse_csr()
    for fn in (lambda: ops.score_dense(U, W), lambda: ops.fused_ce(U, W, lab), lambda: ops.topk_eval(U, W, 4),
               lambda: ops.gather_rows(W, lab), lambda: ops.gather_dot(U, W, idx), lambda: ops.normalize_rows(W),
               lambda: ops.spmm(A, W), lambda: ops.ce_backward(U, W, lab, torch.zeros(8), 1.0)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_compute_entry_reports_no_device(lib):
    rc = lib.rb_gather_rows(None, None, None, 0, 1, 8, 0, None)
    assert rc != 0
    assert b"no CUDA device" in lib.rb_last_error() or b"CUDA" in lib.rb_last_error()


def test_source_never_imports_oracle():
    for p in (ROOT / "recboard_b200").rglob("*.py"):
        txt = p.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt, p
