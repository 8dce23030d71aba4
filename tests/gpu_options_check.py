"""Round-2 GPU check of the options that are OFF by default and have only run on the CPU model so far (NOT collected by pytest on
purpose; one option per process, each under its own `timeout`, so that a kernel that hangs or faults costs its own step only):

    for o in tc_b0 b2_skipbuf kc_stage lvc_pipe pipe_rows; do timeout 150 python tests/gpu_options_check.py $o; done

Per option: parity at small shapes (eps against the oracle and against the default path; bitwise where the option promises the same bits),
then the per-class kernel times at config 2 (B = 8, T' = 861) with the option off and on.  `--emu` runs the same script on the CPU
emulation build with a tiny timing shape (a dry run of the script itself)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import fastdiff_b200 as fb  # noqa: E402
from fastdiff_b200.synthetic import make_inputs, make_state_dict  # noqa: E402
from oracle import fastdiff_oracle as O  # noqa: E402

VALUES = {"tc_b0": (1, 2), "b2_skipbuf": (1,), "kc_stage": (1,), "lvc_pipe": (1,), "pipe_rows": (1,)}   # pipe_rows = lvc_pipe + b2_skipbuf
BITWISE = {"tc_b0": False, "b2_skipbuf": True, "kc_stage": True, "lvc_pipe": True, "pipe_rows": True}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    emu = "--emu" in sys.argv
    if len(args) != 1 or args[0] not in VALUES:
        print(__doc__)
        return 2
    opt = args[0]

    def set_opt(eng, v):
        for key in (("lvc_pipe", "b2_skipbuf") if opt == "pipe_rows" else (opt,)):
            eng.set_option(key, v)
    dev = torch.device("cpu" if emu else "cuda:0")
    sd = make_state_dict(1234, g_jitter=0.1)
    W = O.fold_weight_norm(sd)
    net = fb.FastDiff().to(dev).eval()
    if emu:
        import __graft_entry__ as g
        net._lib_path = g.build_emu()
    net.load_state_dict(sd)
    net.mode = "tc_3xf16"
    sync = (lambda: None) if emu else torch.cuda.synchronize
    ok = True
    for B, Tm in ([(2, 9)] if emu else [(1, 5), (2, 33), (2, 129), (3, 17)]):
        x, mel = make_inputs(B, Tm, 21)
        t = torch.tensor([7.413235, 498.0537, 74.99228][:B]).reshape(B, 1)
        ref, inter = O.denoise(W, x, mel, t, return_intermediates=True)
        xd, md, td = x.to(dev), mel.to(dev), t.to(dev)
        e0 = net((xd, md, td)).cpu()
        eng = net.engine()
        for v in VALUES[opt]:
            set_opt(eng, v)
            e1 = net((xd, md, td)).cpu()
            eng.set_option("stop_after", 3)
            net((xd, md, td))
            l0 = eng.debug_read("lvc0", B, Tm).reshape(B, 32, Tm * 8).cpu()
            eng.set_option("stop_after", 99)
            set_opt(eng, 0)
            res = {"eps_default_vs_oracle": (e0 - ref).abs().max().item(), "eps_option_vs_oracle": (e1 - ref).abs().max().item(),
                   "option_vs_default": (e1 - e0).abs().max().item(), "bitwise": bool(torch.equal(e0, e1)),
                   "lvc0_vs_oracle": (l0 - inter["lvc0"]).abs().max().item()}
            print(f"{opt}={v} B={B} T'={Tm}", json.dumps(res), flush=True)
            ok &= res["eps_option_vs_oracle"] < 5e-5 and res["lvc0_vs_oracle"] < 1e-4 and (res["bitwise"] or not BITWISE[opt])
    B, Tm = (1, 4) if emu else (8, 861)
    x, mel = make_inputs(B, Tm, 1)
    x, mel = x.to(dev), mel.to(dev)
    t = torch.full((B, 1), 74.99228, device=dev)
    eng = net.engine()
    n_rep = 1 if emu else 5
    for v in (0,) + VALUES[opt]:
        set_opt(eng, v)
        for _ in range(0 if emu else 3):
            net((x, mel, t))
        sync()
        eng.timing_enable(True)
        for _ in range(n_rep):
            net((x, mel, t))
        sync()
        rep = eng.timing_report()
        eng.timing_enable(False)
        print(f"{opt}={v} kernel ms per evaluation:", json.dumps({k: round(val["ms"] / n_rep, 4) for k, val in rep.items() if val["n"]}), flush=True)
    set_opt(eng, 0)
    print(opt, "PARITY", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
