import sys, os
sys.path.insert(0, os.getcwd())
import torch
import secp256k1_zkp_b200 as b200
dev = torch.device("cuda", 0)
eng = b200.Engine(device=0)
st = torch.cuda.current_stream().cuda_stream
n = 1 << 20
gen = torch.Generator(device=dev); gen.manual_seed(1)
sk = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen); sk[:, 0] &= 0x7F
sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen)
z64, z32 = torch.zeros((n, 64), dtype=torch.uint8, device=dev), torch.zeros((n, 32), dtype=torch.uint8, device=dev)
pts = torch.empty((n, 64), dtype=torch.uint8, device=dev)
eng.ecmult_dev(pts.data_ptr(), z64.data_ptr(), z32.data_ptr(), sk.data_ptr(), n, st)
wsb = eng.ecmult_multi_workspace_bytes(n); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
out = torch.zeros(96, dtype=torch.uint8, device=dev)
shard = int(sys.argv[1]); shards = int(sys.argv[2])
for _ in range(3):
    eng.ecmult_multi_window_shard_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), n, shard, shards, ws.data_ptr(), wsb, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.ecmult_multi_window_shard_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), n, shard, shards, ws.data_ptr(), wsb, st)
e1.record(); torch.cuda.synchronize()
print("shard %d/%d: %.3f ms" % (shard, shards, e0.elapsed_time(e1) / 10))
