"""Round-2 GPU check of the experimental options tc_b0 (block-0 tensor-core LVC kernel) and b2_skipbuf (block 2 with skip rows from memory) ( NOT collected by pytest on purpose: the kernel
has only run on the CPU model so far -- run this under `timeout`):   timeout 120 python tests/gpu_options_check.py
Parity of block 0 / eps against the oracle and the default path at small shapes, then the per-class kernel times at config 2."""
import json
import sys

import torch

sys.path.insert(0, ".")
import fastdiff_b200 as fb  # noqa: E402
from fastdiff_b200.synthetic import make_inputs, make_state_dict  # noqa: E402
from oracle import fastdiff_oracle as O  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    sd = make_state_dict(1234, g_jitter=0.1)
    W = O.fold_weight_norm(sd)
    net = fb.FastDiff().to(dev).eval()
    net.load_state_dict(sd)
    ok = True
    for B, Tm in [(1, 5), (2, 33), (1, 129), (3, 17)]:
        x, mel = make_inputs(B, Tm, 21)
        t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B, 1)
        ref, inter = O.denoise(W, x, mel, t, return_intermediates=True)
        eng = net.engine()
        eng.set_option("tc_b0", 0)
        e0 = net((x.to(dev), mel.to(dev), t.to(dev))).cpu()
        eng.set_option("tc_b0", 2)
        e2 = net((x.to(dev), mel.to(dev), t.to(dev))).cpu()
        eng.set_option("tc_b0", 1)
        e1 = net((x.to(dev), mel.to(dev), t.to(dev))).cpu()
        eng.set_option("stop_after", 3)
        net((x.to(dev), mel.to(dev), t.to(dev)))
        l0 = eng.debug_read("lvc0", B, Tm).reshape(B, 32, Tm * 8).cpu()
        eng.set_option("stop_after", 99)
        errs = {"eps_default": (e0 - ref).abs().max().item(), "eps_tc_b0": (e1 - ref).abs().max().item(), "v1_equals_v2": bool(torch.equal(e1, e2)),
                "lvc0_tc_b0": (l0 - inter["lvc0"]).abs().max().item()}
        print(B, Tm, errs)
        ok &= errs["eps_tc_b0"] < 5e-5 and errs["lvc0_tc_b0"] < 1e-4
    B, Tm = 8, 861
    x, mel = make_inputs(B, Tm, 1)
    x, mel = x.to(dev), mel.to(dev)
    t = torch.full((B, 1), 74.99228, device=dev)
    eng = net.engine()
    for opt in (0, 1, 2):
        eng.set_option("tc_b0", opt)
        for _ in range(3):
            net((x, mel, t))
        torch.cuda.synchronize()
        eng.timing_enable(True)
        for _ in range(5):
            net((x, mel, t))
        torch.cuda.synchronize()
        rep = eng.timing_report()
        eng.timing_enable(False)
        print("tc_b0 =", opt, json.dumps({k: round(v["ms"] / 5, 4) for k, v in rep.items() if v["n"]}))
    # option b2_skipbuf: LVC block 2 with skip rows from memory -- must give the default's bits; per-class times
    x2, mel2 = make_inputs(2, 33, 5)
    t2 = torch.tensor([[7.413235], [498.0537]], device=dev)
    eng.set_option("tc_b0", 0)
    eng.set_option("b2_skipbuf", 0)
    ea = net((x2.to(dev), mel2.to(dev), t2))
    eng.set_option("b2_skipbuf", 1)
    eb = net((x2.to(dev), mel2.to(dev), t2))
    same = bool(torch.equal(ea, eb))
    print("b2_skipbuf bitwise equal to default:", same)
    ok &= same
    for opt in (0, 1):
        eng.set_option("b2_skipbuf", opt)
        for _ in range(3):
            net((x, mel, t))
        torch.cuda.synchronize()
        eng.timing_enable(True)
        for _ in range(5):
            net((x, mel, t))
        torch.cuda.synchronize()
        rep = eng.timing_report()
        eng.timing_enable(False)
        print("b2_skipbuf =", opt, json.dumps({k: round(v["ms"] / 5, 4) for k, v in rep.items() if v["n"]}))
    # option kc_stage: GEMM epilogue through shared memory + bulk stores -- the default's bits; kc_gemm class time
    eng.set_option("b2_skipbuf", 0)
    eng.set_option("kc_stage", 1)
    ec = net((x2.to(dev), mel2.to(dev), t2))
    same = bool(torch.equal(ea, ec))
    print("kc_stage bitwise equal to default:", same)
    ok &= same
    for opt in (0, 1):
        eng.set_option("kc_stage", opt)
        for _ in range(3):
            net((x, mel, t))
        torch.cuda.synchronize()
        eng.timing_enable(True)
        for _ in range(5):
            net((x, mel, t))
        torch.cuda.synchronize()
        rep = eng.timing_report()
        eng.timing_enable(False)
        print("kc_stage =", opt, json.dumps({k: round(v["ms"] / 5, 4) for k, v in rep.items() if v["n"]}))
    eng.set_option("kc_stage", 0)
    print("PARITY", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
