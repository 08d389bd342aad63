"""Kernel-time breakdown (torch.profiler) of one fused sparse step of the reference's full-size SD U-Net / GauGAN generator.
Development aid, GPU only:  python tools/profile_consumer.py sd|gaugan"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "baseline"))
import torch
import consumers
from sige.utils import dilate_mask, downsample_mask

which = sys.argv[1]
dev = torch.device("cuda", 0)
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
net = (consumers.build_sd("full") if which == "sd" else consumers.build_gaugan("full")).to(dev)
fused = lambda n: n.set_fused(True, dtype=torch.float16, use_graph=False)
if which == "sd":
    consumers.run_sd(net, downsample_mask, device=dev, fused=fused, size="full")
    a = [v.to(dev) for v in consumers.sd_inputs("full")]; args = (a[1], a[3], a[4])
else:
    consumers.run_gaugan(net, downsample_mask, dilate_mask, device=dev, fused=fused, size="full")
    args = (consumers.gaugan_inputs("full")[1].to(dev),)
step = net.fused_step
print("fused launches:", len(step.fused), "eager nodes:", len(step.eager_nodes))
with torch.no_grad():
    for _ in range(3):
        net(*args)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        net(*args)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
