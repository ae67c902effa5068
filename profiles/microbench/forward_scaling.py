"""Time rsf_forward_kernel (SSE only) for a range of batch sizes and block sizes."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
pkg = importlib.import_module("bayesian-markov-chain-monte-carlo_b200")
m = pkg.RateStateModel()
m.Dc = 1325.0
_, acc, _ = m.evaluate()
rng = np.random.default_rng(0)
MODES = os.environ.get("MODES", "parity").split(",")
for blk in (os.environ.get("BLOCKS", "32,64,128").split(",")):
    m.block_threads = int(blk)
    for mode in MODES:
      m.integ_mode = mode
      for c in (32, 64, 128, 256, 512, 1024, 2048, 4096, 4736, 9472, 18944, 37888, 65536, 131072, 262144):
       for _once in (0,):
        dc = torch.from_numpy(rng.uniform(800, 2000, c)).cuda()
        m.evaluate_batch(dc, data=acc, want_acc=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 3
        for _ in range(reps):
            m.evaluate_batch(dc, data=acc, want_acc=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"block {blk:>3s} {mode:6s} C {c:7d}: {ms:8.3f} ms  {c / ms * 1e3:12.0f} solves/s", flush=True)
