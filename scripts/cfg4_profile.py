"""One N=1000 forward + backward (BASELINE config 4) for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segnn_b200 as S
import bench

torch.manual_seed(0)
dev = torch.device("cuda", 0)
N4, H4 = 1000, 128
m4 = S.SEGNN(hidden_features=H4, num_layers=6, lmax_h=1).to(dev).train()
p4, v4, c4 = bench.synthetic_system(1, N4, seed=5)
g = S.GraphBatch(pos=p4.reshape(-1, 3).to(dev), vel=v4.reshape(-1, 3).to(dev), mass=c4.reshape(-1, 1).to(dev),
                 num_graphs=1, n_nodes=N4)
y4 = torch.randn(N4, 6, device=dev)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    m4.zero_grad(set_to_none=True)
    loss = S.target_common_loss(m4(g), y4)
    loss.backward()
    torch.cuda.synchronize()
    print(i, float(loss))
