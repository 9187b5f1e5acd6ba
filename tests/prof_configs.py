"""Timings of the per-GPU shards of BASELINE configs[3] and configs[4] (CUDA events, one GPU):
  config 4: HSTU-style retrieval, 10M-item table / 8 GPUs = 1.25M x 256 bf16 per GPU, top-100, B in {256, 4096}
  config 5: BERT4Rec masked-item CE, 50M items / 8 GPUs = 6.25M x 128 bf16 per GPU + bias, M = 4096 rows"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from recboard_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(4)
PEAK_TF, PEAK_GB = 1687.9, 6545.0
try:
    pk = json.load(open(Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"))
    PEAK_TF, PEAK_GB = pk["bf16_tflops"], pk["hbm_gbs"]
except Exception:
    pass
out = {"peaks": {"bf16_tflops": PEAK_TF, "hbm_gbs": PEAK_GB}}


def t(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


N4, d4, K4 = 1_250_000, 256, 100
W4 = ops.normalize_rows(torch.randn(N4, d4, device=dev, generator=g), out_dtype=torch.bfloat16)
for B in (256, 4096):
    U4 = ops.normalize_rows(torch.randn(B, d4, device=dev, generator=g), out_dtype=torch.bfloat16)
    crow, col = synth.seen_csr(B, N4, g, dev)
    ms = t(lambda: ops.topk_eval(U4, W4, K4, crow, col))
    flop = 2.0 * B * N4 * d4
    nbytes = N4 * d4 * 2 + B * d4 * 2 + col.numel() * 8 + B * K4 * 8
    out[f"config4 top-100 B={B} (1.25M x 256 shard)"] = {
        "ms": round(ms, 4), "pairs_per_s": B * N4 / ms * 1e3, "algorithmic_tflops": round(flop / ms / 1e9, 1),
        "frac_of_tensor_peak": round(flop / ms / 1e9 / PEAK_TF, 3), "algorithmic_GBps": round(nbytes / ms / 1e6, 1),
        "frac_of_hbm_peak": round(nbytes / ms / 1e6 / PEAK_GB, 3)}
del W4

M5, N5, d5 = 4096, 6_250_002, 128
U5 = synth.embeddings(M5, d5, g, dev, torch.bfloat16)
W5 = synth.embeddings(N5, d5, g, dev, torch.bfloat16)
bias = torch.randn(N5, device=dev, generator=g) * 0.2
lab = synth.zipf_ids(M5, N5 - 2, g, dev) + 2
m, l, ll = ops.ce_rowstats(U5, W5, lab, bias=bias)
lse = m + torch.log(l)
t_f = t(lambda: ops.ce_rowstats(U5, W5, lab, bias=bias, want_dU=True), 3)
t_b = t(lambda: ops.ce_backward(U5, W5, lab, lse, 1.0 / M5, bias=bias, need_dU=False, need_dW=True, need_dbias=True), 3)
flop = 2.0 * M5 * N5 * d5
out["config5 CE train M=4096 (6.25M x 128 shard + bias)"] = {
    "fwd_dU_ms": round(t_f, 3), "dW_dbias_ms": round(t_b, 3), "pairs_per_s": M5 * N5 / (t_f + t_b) * 1e3,
    "algorithmic_tflops": round(3 * flop / (t_f + t_b) / 1e9, 1), "frac_of_tensor_peak": round(3 * flop / (t_f + t_b) / 1e9 / PEAK_TF, 3),
    "peak_mem_GiB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}
print(json.dumps(out, indent=1))
