mkdir -p gpurun_out/r2x19
python - <<'PY' > gpurun_out/r2x19/crf_chunks.txt 2>&1
import sys, time, os, subprocess
for c in (8, 16, 32, 64):
    out = subprocess.run([sys.executable, "-c", """
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from digipathai_b200 import engine
rng = np.random.default_rng(0)
n = 64
rgb = torch.from_numpy(rng.integers(0, 256, (n, 256, 256, 3)).astype(np.uint8)).cuda()
p1 = torch.from_numpy(rng.uniform(0, 1, (n, 256, 256)).astype(np.float32)).cuda()
lab = engine.dense_crf(rgb, p1); torch.cuda.synchronize()
t = time.time(); lab = engine.dense_crf(rgb, p1); torch.cuda.synchronize(); dt = time.time() - t
import hashlib
print(f'{dt*1e3:.1f} ms for 64 tiles = {dt*1e3/n:.3f} ms per tile, labels sha {hashlib.sha256(lab.cpu().numpy().tobytes()).hexdigest()[:12]}')
"""], env=dict(os.environ, DP_CRF_CHUNK=str(c)), capture_output=True, text=True)
    print("chunk", c, out.stdout.strip(), out.stderr.strip()[-300:])
PY
cat gpurun_out/r2x19/crf_chunks.txt
timeout 300 python -m pytest tests/test_gpu_crf.py tests/test_gpu_utils_crf.py -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload config5 --slide 8192 --steps 2 > gpurun_out/r2x19/config5.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2x19/config5.json').read().strip().splitlines()[-1])['config5']; print(d['tiles_per_s'], d['loop_ms'], d['crf_ms'])"
