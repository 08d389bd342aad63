"""Per-stage timeline (globaltimer) of every tcgen05 fused launch of the DDPM step engine (development aid)."""
import ctypes, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sige_b200 import _cabi, ops
from sige_b200.engine import DDPMStepEngine
from sige_b200.masks import downsample_mask
from sige_b200.workloads.ddpm import DDPMConfig, SIGEDDPMUNet, init_deterministic, synthetic_inputs

dev = torch.device("cuda", 0)
cfg = DDPMConfig()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    model = init_deterministic(SIGEDDPMUNet(cfg), seed=0).eval().to(dev).half().to(memory_format=torch.channels_last)
x0, x1, mask, t = synthetic_inputs(cfg, 0.012, seed=0)
cl = lambda a: a.to(dev).half().contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    model.set_mode("full"); model(cl(x0), t.to(dev))
    model.set_masks(downsample_mask(mask.to(dev), min_res=8)); model.set_mode("sparse")
pdl = "--pdl" in sys.argv
eng = DDPMStepEngine(model, cl(x1), use_graph=False, tc5=True, pdl=pdl)
lib = _cabi.lib()
lib.sige_debug_set_trace.argtypes = [ctypes.c_void_p]
buf = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
names = ["start", "setup", "idx", "ldg", "store", "mmaA", "mmaEnd", "accRdy", "tmemLd", "clSync", "stored", "end"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
tot = 0.0
rows = []
for f in eng.fused:
    buf.zero_(); flush.fill_(1); torch.cuda.synchronize()
    lib.sige_debug_set_trace(buf.data_ptr())
    f.launch(s)
    torch.cuda.synchronize()
    lib.sige_debug_set_trace(None)
    tt = buf.view(-1, 16).cpu()
    tt = tt[tt[:, 0] > 0]
    if tt.shape[0] == 0:
        rows.append((f.name, None)); continue
    t0 = tt[:, 0].min()
    rel = (tt - t0).float(); rel[tt == 0] = float("nan")
    med = [float(torch.nanmedian(rel[:, i])) for i in range(12)]
    end = float(rel[:, 11][~torch.isnan(rel[:, 11])].max())
    tot += end
    rows.append((f.name, (tt.shape[0], f.desc.ksplit, end, med)))
for name, r in rows:
    if r is None:
        print("%-28s (mma.sync kernel, not traced)" % name); continue
    n, ks, end, med = r
    print("%-28s ctas %3d end %6.0f | " % (name, n, end) + " ".join("%s %.0f" % (a, b) for a, b in zip(names[1:], med[1:])))
print("sum of traced kernel spans: %.1f us over %d launches" % (tot / 1e3, len(rows)))
