"""Device Fisher (as_fisher_tests_host) vs the scalar host form on call tables of the c5 shape (50,000x, 0.5-1 % VAF
somatic calls plus germline calls); run under gpurun."""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from amplisolve_b200 import Context, fisher_test  # noqa: E402

rng = np.random.default_rng(5)
n = 200_000
depth = np.maximum(2000, (50000 * rng.lognormal(0, 0.5, n))).astype(np.int64)
fw = rng.binomial(depth, 0.5)
bw = depth - fw
germ = rng.random(n) < 0.25
vaf = np.where(germ, rng.choice([0.5, 1.0], n), rng.uniform(0.005, 0.01, n))
tables = np.stack([fw, bw, rng.binomial(fw, vaf), rng.binomial(bw, vaf)], 1).astype(np.int32)
with Context(0) as ctx:
    ctx.fisher_tests(tables[:1000])
    t0 = time.perf_counter()
    p = ctx.fisher_tests(tables)
    t_dev = time.perf_counter() - t0
m = 300
t0 = time.perf_counter()
host = np.array([fisher_test(*t) for t in tables[:m].tolist()])
t_host = (time.perf_counter() - t0) / m * n
ok = host > 1e-290
rel = float(np.max(np.abs(p[:m] - host)[ok] / host[ok])) if ok.any() else 0.0
terms = float(np.minimum(tables[:, 0] + tables[:, 2], tables[:, 2] + tables[:, 3]).sum())
print(json.dumps({"tables": n, "pdf_terms": terms, "device_s": t_dev, "host_scalar_s_extrapolated_1_thread": t_host,
                  "max_rel_diff_on_sample": rel, "terms_per_s_device": terms / t_dev}))
