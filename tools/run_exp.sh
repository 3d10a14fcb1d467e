timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "search or tokenize_host or degenerate" 2>&1 | tail -3
cat > /tmp/t16.py <<'PY'
import sys, os, torch
sys.path.insert(0, "/root/repo")
import gesture2vec_b200 as g
from gesture2vec_b200 import _lib
N, K, D = 1000000, int(sys.argv[1]), 400
dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[sys.argv[2]]
dev = torch.device("cuda:0")
z = torch.randn(N, D, device=dev).to(dt); E = torch.randn(K, D, device=dev)
cb = g.prepare_codebook(E)
for fl, nm in ((_lib.ALGO_TC, "full"), (_lib.ALGO_TC | _lib.NO_RECHECK, "fast")):
    for _ in range(3): g.vq_search(z, E, cb, flags=fl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.vq_search(z, E, cb, flags=fl)
    e1.record(); torch.cuda.synchronize()
    print(f"K={K} {sys.argv[2]} tmem16={os.environ.get('G2V_TC_TMEM16','1')} {nm}: {e0.elapsed_time(e1)/5:.3f} ms")
PY
for t in 1 0; do G2V_TC_TMEM16=$t timeout 30 python /tmp/t16.py 512 bf16; done
timeout 30 python /tmp/t16.py 400 f16
timeout 30 python /tmp/t16.py 1024 f32
timeout 30 python tools/tc_time.py 262144 16384 400 fast | tail -1
timeout 30 python tools/tc_time.py 262144 4096 400 fast | tail -1
