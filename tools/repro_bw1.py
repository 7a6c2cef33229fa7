import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from animatable_nerf_b200 import config, synthetic
from animatable_nerf_b200.tpose_nerf_network import Network
dev = torch.device('cuda:0')
sd = synthetic.make_state_dict(seed=0)
net = Network(config.make_cfg(b200_bw_precision=1)); net.load_state_dict(sd); net = net.to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4101
g = torch.Generator().manual_seed(1)
pts = torch.rand(1, n, 3, generator=g).to(dev)
init = torch.softmax(torch.randn(1, 24, n, generator=g), dim=1).to(dev)
out = net.calculate_neural_blend_weights(pts, init, torch.tensor([1]))
torch.cuda.synchronize()
print('bw x1 ok', out.shape, float(out.sum()))
