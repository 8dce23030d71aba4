"""Quick hardware check of the piece-row LVC path (NOT collected by pytest; run first in a GPU call so that a protocol bug costs seconds):
eps / lvc1 / lvc2 against the oracle and against the k_lvc_layer_h path at small and ragged shapes, then per-class kernel times at config 2."""
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
    net.mode = "tc_3xf16"
    ok = True
    for B, Tm in [(1, 3), (1, 5), (2, 9), (3, 1), (2, 33), (1, 86), (2, 129), (1, 861)]:
        x, mel = make_inputs(B, Tm, 21)
        t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B, 1)
        ref, inter = O.denoise(W, x, mel, t, return_intermediates=True)
        xd, md, td = x.to(dev), mel.to(dev), t.to(dev)
        eng = net.engine()
        res = {}
        outs = {}
        for lp in (1, 0):
            eng.set_option("lvc_p", lp)
            e = net((xd, md, td)).cpu()
            outs[lp] = e
            l2 = eng.debug_read("lvc2", B, Tm).reshape(B, 32, Tm * 256).cpu()
            eng.set_option("stop_after", 4)
            net((xd, md, td))
            l1 = eng.debug_read("lvc1", B, Tm).reshape(B, 32, Tm * 64).cpu()
            eng.set_option("stop_after", 99)
            res[f"p{lp}"] = {"eps": (e - ref).abs().max().item(), "lvc1": (l1 - inter["lvc1"]).abs().max().item(),
                            "lvc2": (l2 - inter["lvc2"]).abs().max().item()}
        res["p_vs_h"] = (outs[1] - outs[0]).abs().max().item()
        e2 = net((xd, md, td)).cpu()   # lvc_p is 0 here; determinism of the new path checked next
        eng.set_option("lvc_p", 1)
        a, b = net((xd, md, td)).cpu(), net((xd, md, td)).cpu()
        res["deterministic"] = bool(torch.equal(a, b) and torch.equal(a, outs[1]) and torch.equal(e2, outs[0]))
        print(f"B={B} T'={Tm}", json.dumps(res), flush=True)
        ok &= res["p1"]["eps"] < 5e-5 and res["p1"]["lvc1"] < 2e-4 and res["p1"]["lvc2"] < 2e-4 and res["deterministic"]
    print("saturation flag:", eng.check_saturation())
    B, Tm = 8, 861
    x, mel = make_inputs(B, Tm, 1)
    x, mel = x.to(dev), mel.to(dev)
    t = torch.full((B, 1), 74.99228, device=dev)
    eng = net.engine()
    for lp in (0, 1):
        eng.set_option("lvc_p", lp)
        for _ in range(3):
            net((x, mel, t))
        torch.cuda.synchronize()
        eng.timing_enable(True)
        for _ in range(5):
            net((x, mel, t))
        torch.cuda.synchronize()
        rep = eng.timing_report()
        eng.timing_enable(False)
        print(f"lvc_p={lp} kernel ms per evaluation:", json.dumps({k: round(val["ms"] / 5, 4) for k, val in rep.items() if val["n"]}), flush=True)
    print("PARITY", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
