#!/usr/bin/env python
"""Regenerate profiles/sass/r02_*.sass (one listing per default-path kernel, encodings stripped) and r02_mnemonics.txt from the in-tree
library:   python tools/dump_sass.py        (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastdiff_b200", "csrc", "libfastdiff_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
KERNELS = {   # file stem -> mangled name
    "k_dblock0_tc": "_ZN2fd12k_dblock0_tcENS_10DbTcParamsEPKfPfiiii",
    "k_dblock8": "_ZN2fd8k_dblockILi8ELb0EEEvNS_8DbParamsEPKfPfii",
    "k_embed": "_ZN2fd7k_embedENS_11EmbedParamsEPKfNS_10EmbedStepsEPfS4_i",
    "k_fill_normal": "_ZN2fd13k_fill_normalEPfimmjNS_8NoiseWinEPKy",
    "k_final": "_ZN2fd7k_finalENS_11FinalParamsEPKfS2_S2_PfS3_i",
    "k_kc_gemm_tc2_f16_b0p": "_ZN2fd13k_kc_gemm_tc2ILb1ELi16ELb1ELb0EEEvNS_7KcgMapsEPKfS3_S3_Pfiiifffi",
    "k_kp_hidden_tc": "_ZN2fd14k_kp_hidden_tcENS_10KpTcParamsEPKfS2_PfS3_S3_ii",
    "k_lvc_layer_b0h": "_ZN2fd15k_lvc_layer_b0hENS_10LvcHParamsEPKfS2_S2_Pfiiiiffii",
    "k_lvc_p_hop256": "_ZN2fd7k_lvc_pILi256EEEvNS_10LvcPParamsE",
    "k_lvc_p_hop64": "_ZN2fd7k_lvc_pILi64EEEvNS_10LvcPParamsE",
    "k_upsample8_simt": "_ZN2fd10k_upsampleILi8EEEvPKfS2_S2_Pfi",
    "k_upsample_p4": "_ZN2fd13k_upsample_p4ENS_9Up4ParamsE",
    "k_upsample_tc_r8_pout": "_ZN2fd13k_upsample_tcILi8ELb1EEEvPKfS2_S2_S2_PfiiiNS_6UpPOutE",
    "k_zero_pads": "_ZN2fd11k_zero_padsEPfii",
}
TC = ("LDTM", "UBLKCP", "UBLKPF", "UTCATOMSWS", "UTCBAR", "UTCHMMA", "UTMALDG")
SIMT = ("FFMA", "MUFU", "F2FP", "STG", "LDG", "STS", "LDS")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    by_name = {p.split("\n", 1)[0].strip(): p for p in parts[1:]}
    rows = []
    for stem, mangled in KERNELS.items():
        body = by_name[mangled]
        ins = []
        for line in body.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", line)
            if m:
                ins.append((m.group(1), m.group(2).strip()))
        with open(os.path.join(OUT, f"r02_{stem}.sass"), "w") as f:
            f.write(f".text.{mangled}:\n")
            for addr, text in ins:
                f.write(f"        /*{addr}*/                   {text} ;\n")
        ops = collections.Counter()
        for _, text in ins:
            op = text.split()[1] if text.startswith("@") else text.split()[0]
            ops[op] += 1
        tc = ", ".join(f"{k} {v}" for k, v in sorted(ops.items()) if k.startswith(TC))
        simt = ", ".join(f"{p} {sum(v for k, v in ops.items() if k.startswith(p))}" for p in SIMT)
        ef = sum(v for k, v in ops.items() if k.startswith("STG") and ".EF" in k)
        rows.append(f"{stem:30s} {len(ins):5d} SASS instructions | {tc} | {simt}" + (f", of the STG {ef} evict-first" if ef else ""))
    with open(os.path.join(OUT, "r02_mnemonics.txt"), "w") as f:
        f.write("# mnemonic counts of the default-path kernels, `cuobjdump -sass` of the committed build (round 2, final state; tools/dump_sass.py)\n")
        f.write("\n".join(rows) + "\n")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
