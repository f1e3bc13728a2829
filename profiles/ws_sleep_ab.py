"""A/B of the forward ws kernel's mbarrier poll back-off (MDL_WS_SLEEP ns) on the roofline workload."""
import os, sys, subprocess
for ns in ("0", "20", "50", "100", "200", "500"):
    env = dict(os.environ, MDL_WS_SLEEP=ns, MDL_CGCONV_IMPL="ws")
    out = subprocess.run([sys.executable, "bench.py", "--roofline-only"], env=env, capture_output=True, text=True).stdout
    import json
    d = json.loads(out.strip().splitlines()[-1])
    print(f"sleep {ns:>4s} ns: fwd {d['fwd']['ms']:.3f} ms  bwd {d['bwd_both_passes']['ms']:.3f} ms", flush=True)
