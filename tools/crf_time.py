import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from digipathai_b200 import engine
rng = np.random.default_rng(0)
for n in (1, 8, 32):
    rgb = torch.from_numpy(rng.integers(0, 256, (n, 256, 256, 3)).astype(np.uint8)).cuda()
    p1 = torch.from_numpy(rng.uniform(0, 1, (n, 256, 256)).astype(np.float32)).cuda()
    engine.dense_crf(rgb, p1); torch.cuda.synchronize()
    t = time.time(); engine.dense_crf(rgb, p1); torch.cuda.synchronize(); dt = time.time() - t
    print(f"dense_crf {n} tiles 256x256, 10 iterations: {dt*1e3:.1f} ms ({dt*1e3/n:.1f} ms per tile)")
