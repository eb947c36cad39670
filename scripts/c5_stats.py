"""How the c5-shaped workload (ultra-deep, low VAF) loads the caller's stages: candidates, m-screen survivors, calls."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from amplisolve_b200 import Context, calls_from_device  # noqa: E402

P, S, T = 4000, 100, 2000
depth = float(sys.argv[1]) if len(sys.argv) > 1 else 50000.0
som, vaf = (0.0005, (0.005, 0.01)) if depth > 10000 else (2e-4, (0.01, 0.2))
with Context(0) as ctx:
    gen = dict(seed=20183, mean_depth=depth, twin_period=6)
    normals, ref = ctx.synth_counts_dev(S, P, **gen)
    tumours, _ = ctx.synth_counts_dev(T, P, somatic_rate=som, vaf=vaf, sample_offset=1 << 20, want_ref=False, **gen)
    nxt, head = ctx.synth_twin_links_dev(P, seed=20183, twin_period=6)
    out = ctx.alloc_noise_outputs(P)
    ctx.estimate_thresholds_dev(normals, 0.002, 100, out, nxt, head)
    view = ctx.thresholds_caller_view_dev(out["thr"])
    calls = torch.zeros(48 * 4_000_000, dtype=torch.uint8, device="cuda")
    n = torch.zeros(1, dtype=torch.int64, device="cuda")
    ctx.call_variants_dev(tumours, ref, view, 100, calls, n)
    ncalls = int(n.item())
    t = tumours.cpu().numpy().view(np.uint32).astype(np.int64)        # [T][2][P][4]
    e = view.cpu().numpy().astype(np.float64)                          # [P][4][2]
    r = ref.cpu().numpy()
fw, bw = t[:, 0], t[:, 1]
FW, BW = fw.sum(-1), bw.sum(-1)
cov = (FW >= 100) & (BW >= 100)
notref = np.ones((P, 4), bool)
notref[np.arange(P), r] = False
cand = cov[..., None] & notref[None] & (fw > 0) & (bw > 0)
eff = np.where(e == 0, 0.0010008, e)
m_fw = FW[..., None] * eff[None, :, :, 0]
m_bw = BW[..., None] * eff[None, :, :, 1]
passm = cand & ~((m_fw >= fw) & (m_fw > 1)) & ~((m_bw >= bw) & (m_bw > 1))
pairs = T * P * 3
print(f"depth {depth:g}: pairs {pairs}, candidates {cand.sum()} ({cand.sum()/pairs:.4f}), pass m-screen {passm.sum()} "
      f"({passm.sum()/pairs:.5f}), calls {ncalls} ({ncalls/pairs:.5f}); calls/survivors {ncalls/max(1,passm.sum()):.3f}")
k = fw[passm]
print("k_fw of survivors: median", np.median(k), "p90", np.percentile(k, 90), " m_fw median", np.median(m_fw[passm]))
