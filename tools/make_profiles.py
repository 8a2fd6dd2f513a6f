"""Turns the captures of tools/final_measure.sh TAG (gpurun_out/) into the tracked files under profiles/:
  profiles/<round>_launches_<tag>.txt        launch list (ns per launch, launch order) + each kernel's share of the summed time
  profiles/<round>_ncu_top_kernels_<tag>.md  ncu --set full summaries (tools/ncu_summary.py) of the sampler and the two tcgen05 kernels
  profiles/<round>_sass_hist_<tag>.md        executed-instruction histogram by SASS opcode class of the three kernels (ncu source page)
                                             + static cuobjdump counts of the tensor-core / TMA mnemonics in the built objects
usage: python tools/make_profiles.py r02q [r02]"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
GO = os.path.join(ROOT, "gpurun_out")
PR = os.path.join(ROOT, "profiles")


def launches():
    src = os.path.join(GO, f"launches_{tag}.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    tot = collections.Counter()
    cnt = collections.Counter()
    lines = []
    for r in rows:
        name, ns = r[4], int(float(r[-1]))
        short = re.sub(r"\(.*", "", name)[:110]
        tot[short] += ns
        cnt[short] += 1
        lines.append(f"{name[:118]:118s} {ns:>10d}")
    total = sum(tot.values())
    with open(os.path.join(PR, f"{rnd}_launches_{tag}.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 1 "
                f"--no-cpu-baseline --no-others (tools/final_measure.sh {tag}); ns per launch in launch order, then the shares\n")
        f.write("\n".join(lines) + "\n\n# share of the summed kernel time (launches, total ns, share)\n")
        for k, v in tot.most_common():
            f.write(f"{k:110s} {cnt[k]:4d} {v:>12d} {100.0 * v / total:6.2f} %\n")


def ncu_summaries():
    out = []
    for part in ("sampler", "tc"):
        rep = os.path.join(GO, f"prof_{tag}_{part}.ncu-rep")
        if os.path.exists(rep):
            out.append(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep],
                                      capture_output=True, text=True).stdout.replace(ROOT + "/", ""))
    open(os.path.join(PR, f"{rnd}_ncu_top_kernels_{tag}.md"), "w").write("\n".join(out))


CLASSES = [("packed FP32 (FFMA2/FMUL2/FADD2)", r"^(FFMA2|FMUL2|FADD2)$"), ("scalar FP32 (FFMA/FMUL/FADD/FMNMX/FSEL/FSETP/FCHK)", r"^(FFMA|FMUL|FADD|FMNMX|FSEL|FSETP|FCHK|FSET)$"),
           ("MUFU", r"^MUFU$"), ("LDS/STS/LDSM", r"^(LDS|STS|LDSM|STSM)$"), ("LDG/STG/LD/ST/ATOM/RED", r"^(LDG|STG|LD|ST|ATOM|ATOMG|RED|LDL|STL|ATOMS)$"),
           ("LDGSTS/UBLKCP/UTMALDG (async copies)", r"^(LDGSTS|UBLKCP|UTMALDG|UTMASTG|LDGDEPBAR|DEPBAR)$"),
           ("IMAD/IADD3/LEA/VIADD (integer arithmetic)", r"^(IMAD|IADD3|LEA|VIADD|IADD|IABS|VIMNMX|IMNMX)$"),
           ("SHF/LOP3/PRMT/SEL/ISETP/PLOP3/MOV (bit and predicate work)", r"^(SHF|LOP3|PRMT|SEL|ISETP|PLOP3|MOV|SGXT|BMSK|POPC|FLO|P2R|R2P|CS2R)$"),
           ("I2F/F2I/I2FP/F2FP (conversions)", r"^(I2F|F2I|I2FP|F2FP|F2F|I2I|FRND)$"), ("REDUX/SHFL/VOTE/MATCH (warp collectives)", r"^(REDUX|SHFL|VOTE|MATCH|VOTEU)$"),
           ("BRA/BSSY/BSYNC/WARPSYNC/EXIT/CALL (control)", r"^(BRA|BSSY|BSYNC|WARPSYNC|EXIT|CALL|RET|BREAK|NANOSLEEP|YIELD|NOP|BRX|JMP)$"),
           ("BAR/SYNCS/MEMBAR/FENCE (barriers, mbarrier)", r"^(BAR|SYNCS|MEMBAR|FENCE|ERRBAR|CCTL|ARRIVES|B2R)$"),
           ("UTCHMMA/UTCBAR/UTCATOMSWS/LDTM/STTM (tcgen05)", r"^(UTCHMMA|UTCQMMA|UTCIMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTCCP|UTCSHIFT)$"),
           ("uniform datapath (U*, R2UR, S2UR, LDC*)", r"^(U[A-Z0-9]+|R2UR|S2UR|S2R|LDC|LDCU|ELECT)$"), ("DFMA/DADD/DMUL/F2F.F64 (FP64)", r"^(DFMA|DADD|DMUL|DSETP)$")]


def sass_hist():
    out = [f"# Executed-instruction histogram by SASS class (`ncu --page source`, `tools/final_measure.sh {tag}`)\n",
           "Per kernel: warp-level instructions executed in one launch at C4, share of all executed instructions, and "
           "the share of the warp-stall samples that landed on those instructions.\n"]
    for part, pat in (("sampler", "sample_ratio"), ("tc", "infid_fwd_tc"), ("tc", "infid_grad_tc")):
        rep = os.path.join(GO, f"prof_{tag}_{part}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"],
                             capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        kname, hdr = rows[0][1], rows[1]
        iS, iI, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        by, smp, ops = collections.Counter(), collections.Counter(), collections.Counter()
        static = 0
        for r in rows[2:]:
            try:
                c, s = int(r[iI]), int(r[iSm])
            except (ValueError, IndexError):
                continue
            t = r[iS].strip().split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            static += 1
            ops[op] += c
            for cname, rx in CLASSES:
                if re.match(rx, op):
                    by[cname] += c
                    smp[cname] += s
                    break
            else:
                by["other: " + op] += c
                smp["other: " + op] += s
        tot, ts = sum(by.values()), max(1, sum(smp.values()))
        out.append(f"\n## `{kname[:120]}`\n\n{static} static SASS instructions, {tot:.4g} executed warp instructions per launch\n")
        out.append("| class | executed | share | stall samples |\n|---|---|---|---|")
        for k, v in by.most_common():
            if v * 1000 >= tot:
                out.append(f"| {k} | {v} | {100.0 * v / tot:.2f} % | {100.0 * smp[k] / ts:.2f} % |")
        out.append("\ntop opcodes: " + ", ".join(f"`{k}` {100.0 * v / tot:.1f} %" for k, v in ops.most_common(14)))
    # static counts of the mnemonics that prove tcgen05 / TMA / packed FP32 in the built objects
    out.append("\n## Static mnemonic counts in the built objects (`cuobjdump -sass netket_fidelity_b200/csrc/build/*.o`)\n")
    out.append("| object | UTCHMMA | LDTM | UTCBAR | UBLKCP | UTMALDG | LDGSTS | SYNCS | FFMA2 | FMUL2 | REDUX |\n|---|---|---|---|---|---|---|---|---|---|---|")
    bdir = os.path.join(ROOT, "netket_fidelity_b200", "csrc", "build")
    for o in ("forward_tc.o", "grad_tc.o", "ratio_f32_L32.o", "sr.o"):
        p = os.path.join(bdir, o)
        if not os.path.exists(p):
            continue
        sass = subprocess.run(["cuobjdump", "-sass", p], capture_output=True, text=True).stdout
        c = collections.Counter(m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", sass, re.M))
        out.append(f"| `{o}` | " + " | ".join(str(c.get(k, 0)) for k in
                                             ("UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "FFMA2", "FMUL2", "REDUX")) + " |")
    open(os.path.join(PR, f"{rnd}_sass_hist_{tag}.md"), "w").write("\n".join(out) + "\n")


launches()
ncu_summaries()
sass_hist()
print("wrote", [f for f in sorted(os.listdir(PR)) if tag in f])
