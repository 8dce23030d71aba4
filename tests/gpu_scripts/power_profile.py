"""Energy per kernel class of one reverse step at config 2 on a B200 (NOT collected by pytest):
    python tests/gpu_scripts/power_profile.py [prefix ...]
The denoiser is run as growing prefixes (`stop_after` = 1 .. 99) in a loop of ~2 s each; NVML's total-energy counter and the wall clock give
joules and milliseconds per iteration, the differences between prefixes the share of each class; SM clocks are sampled while the loop runs."""
import sys
import time

import pynvml
import torch

sys.path.insert(0, ".")
import fastdiff_b200 as fb  # noqa: E402
from fastdiff_b200.synthetic import make_inputs, make_state_dict  # noqa: E402

pynvml.nvmlInit()
dev = pynvml.nvmlDeviceGetHandleByIndex(0)


def energy_mj():
    return pynvml.nvmlDeviceGetTotalEnergyConsumption(dev)


net = fb.FastDiff().cuda().eval()
net.load_state_dict(make_state_dict(1234))
B, Tm = 8, 861
x, mel = make_inputs(B, Tm, 3)
t = torch.full((B, 1), 74.99)
data = (x.cuda(), mel.cuda(), t.cuda())
net(data)
eng = net.engine()
eng.set_option("overlap", 0)   # serial order: a prefix is then exactly the kernels before the cut
torch.cuda.synchronize()
time.sleep(1.0)
e0, t0 = energy_mj(), time.perf_counter()
time.sleep(1.5)
p_idle = (energy_mj() - e0) / 1e3 / (time.perf_counter() - t0)
print(f"idle: {p_idle:.0f} W   power limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(dev) / 1e3:.0f} W")
names = {1: "embed + kernel predictor + kernel_conv GEMM", 2: "+ DBlocks (+ block-0 upsample)", 3: "+ LVC block 0", 4: "+ upsample 1 + LVC block 1",
         5: "+ upsample 2 + LVC block 2", 99: "+ final conv + update (whole reverse step)"}
prev_e = prev_t = 0.0
print("prefix | ms/iter | J/iter | avg W | SM MHz (median while running) | class: ms, J, W")
stops = tuple(int(a) for a in sys.argv[1:]) or (1, 2, 3, 4, 5, 99)   # optional: the prefixes to run (differences are then between those)
for stop in stops:
    eng.set_option("stop_after", stop)
    for _ in range(20):
        net(data)
    torch.cuda.synchronize()
    clocks = []
    n = 0
    e0, t0 = energy_mj(), time.perf_counter()
    while time.perf_counter() - t0 < 2.0:
        for _ in range(20):
            net(data)
        n += 20
        clocks.append(pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM))
        torch.cuda.synchronize()
    t1, e1 = time.perf_counter(), energy_mj()
    ms, joule = (t1 - t0) * 1e3 / n, (e1 - e0) / 1e3 / n
    clocks.sort()
    dms, dj = ms - prev_t, joule - prev_e
    print(f"{stop:3d} {names[stop]:48s} | {ms:7.3f} | {joule:6.3f} | {joule / ms * 1e3:5.0f} | {clocks[len(clocks) // 2]:5d} | {dms:6.3f} ms {dj:6.3f} J {dj / dms * 1e3 if dms > 0 else 0:5.0f} W")
    prev_t, prev_e = ms, joule
eng.set_option("stop_after", 99)
eng.set_option("overlap", 1)
