"""PCIe copy rates on this box: pinned H2D / D2H, 1-D vs 2-D (pitched) copies, split copies, both directions at once."""
import torch
n = 16_000_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(3)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(12_531_600, dtype=torch.uint8).pin_memory()
d2 = torch.empty(12_531_600, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("H2D 1-D 16 MB        %.3f ms" % t(lambda: d.copy_(h, non_blocking=True)))
print("D2H 1-D 12.5 MB      %.3f ms" % t(lambda: h2.copy_(d2, non_blocking=True)))
for parts in (2, 4, 8, 16, 64):
    step = n // parts
    print("H2D 1-D in %2d parts   %.3f ms" % (parts, t(lambda: [d[k*step:(k+1)*step].copy_(h[k*step:(k+1)*step], non_blocking=True) for k in range(parts)])))
hv = h.view(4000, 4000)
for pitch in (4000, 4016, 4096):
    dp = torch.empty(4000, pitch, dtype=torch.uint8, device="cuda")
    print("H2D 2-D pitch %d    %.3f ms" % (pitch, t(lambda: dp[:, :4000].copy_(hv, non_blocking=True))))
hw = h[:15_968_000].view(4000, 3992)
dp = torch.empty(4000, 4000, dtype=torch.uint8, device="cuda")
print("H2D 2-D host pitch 3992 -> 4000  %.3f ms" % t(lambda: dp[:, :3992].copy_(hw, non_blocking=True)))
def both():
    e = torch.cuda.Event(); e.record()
    s1.wait_event(e); s2.wait_event(e)
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
print("H2D 16 MB || D2H 12.5 MB  %.3f ms" % t(both))
