#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): which sm_100a instructions the kernels contain
and what each hot kernel costs in registers / shared memory.

    python tools/sass_evidence.py > profiles/sass_evidence_rNN.txt

Reads deep_recommenders_b200/lib/libdeeprec_b200.so with cuobjdump (CUDA toolkit). The mnemonics are the ones
/opt/skills/guides/B200_PROFILING.md lists as proof of tcgen05 / TMA / TMEM use:
    UTCHMMA     tcgen05.mma (kind::tf32 lowers to the HMMA-class UTC op)
    UTMALDG     cp.async.bulk.tensor (TMA load)        UTMACCTL.PF   tensormap prefetch
    LDTM        tcgen05.ld (TMEM -> registers)         UTCBAR        tcgen05.commit -> mbarrier
    UTCATOMSWS  tcgen05.alloc / dealloc                SYNCS.*       mbarrier arrive / try_wait
    REDG...F32x4  red.global.add.v4.f32 (row scatter-add)   LDG.E.NA.128.CONSTANT  ld.global.nc.L1::no_allocate.v4
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "deep_recommenders_b200", "lib", "libdeeprec_b200.so")

MNEMONICS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "UCGABAR", "STTM", "UTMACCTL", "UTCBAR", "UTCATOMSWS", "LDTM", "SYNCS", "REDG", "MATCH", "LDG.E.NA",
             "ATOMG", "REDUX", "FFMA", "HMMA", "IMMA"]
# the instantiations the C2/C3/C4 steps actually launch (profiles/launches_*.json), not all ~600 template variants
HOT = ["embed_fm_fwd_kernel<4, long, 8, false, 0, true", "gemm_tc_pair_kernel<256, 3, ", "gemm_tc_pair_kernel<128, 4, true, true",
       "gemm_tc_kernel<256, 2, ", "gemm_tc_kernel<32, 4, ", "head_bce_kernel", "actgrad_colsum_v4_kernel<2>",
       "embed_fm_fwd_kernel<8, long, 8, ", "embed_fm_fwd_kernel<8, long, 4, ", "embed_fm_bwd_sp_kernel<8, long, 2, ",
       "embed_fm_bwd_kernel<8, long, 2, true", "gather_rows_kernel<long", "scatter_add_rows_kernel<long",
       "gemm_tc_kernel<128, 3, ", "split_tf32", "sgemm_kernel<128, 128, 16, 8, 8, false, false>",
       "sgemm_kernel<128, 32, 16, 8, 2, false, false>", "sgemm_kernel<256, 16, 16, 8, 2, false, false>",
       "inbatch_softmax_kernel", "actgrad_colsum_kernel", "hard_negative", "bucket", "bce_kernel", "sgd_kernel"]


def run(*cmd):
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
    return out.splitlines()


def main():
    if not os.path.exists(SO):
        sys.exit(f"{SO} missing: run `python __graft_entry__.py` (build) first")
    sass = run("cuobjdump", "-sass", SO)
    per_kernel = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            for k in MNEMONICS:
                if op.startswith(k):
                    per_kernel[cur][op] += 1
    names = list(per_kernel)
    pretty = dict(zip(names, demangle(names)))

    print("# SASS evidence, libdeeprec_b200.so (cuobjdump -sass; arch sm_100a)")
    total = collections.Counter()
    for k in names:
        total.update(per_kernel[k])
    print("\n## library-wide counts of the Blackwell-specific / path-defining opcodes")
    for op, n in sorted(total.items(), key=lambda kv: -kv[1]):
        if not op.startswith("FFMA"):
            print(f"{n:8d}  {op}")

    print("\n## per kernel, the instantiations the C2 / C3 / C4 steps launch")
    seen = set()
    for k in names:
        p = pretty[k]
        fam = next((h for h in HOT if h in p), None)
        if fam is None:
            continue
        short = re.sub(r"\(.*", "", p).replace("void ", "")
        if short in seen:
            continue
        seen.add(short)
        ops = {op: n for op, n in per_kernel[k].items()}
        ffma = sum(n for op, n in ops.items() if op.startswith("FFMA"))
        rest = ", ".join(f"{op}x{n}" for op, n in sorted(ops.items()) if not op.startswith("FFMA"))
        print(f"{short}\n      FFMA x{ffma}; {rest}")

    print("\n## registers / static shared memory (cuobjdump -res-usage); dynamic smem is set at launch")
    res = run("cuobjdump", "-res-usage", SO).splitlines()
    rows = []
    for i, line in enumerate(res):
        m = re.match(r"\s*Function (\S+):", line)
        if m and i + 1 < len(res):
            r = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res[i + 1])
            if r:
                rows.append((m.group(1), *map(int, r.groups())))
    dn = demangle([r[0] for r in rows])
    seen = set()
    for (mangled, reg, stack, sh, loc), p in zip(rows, dn):
        if not any(h in p for h in HOT):
            continue
        short = re.sub(r"\(.*", "", p).replace("void ", "")
        if short in seen:
            continue
        seen.add(short)
        print(f"REG {reg:3d}  STACK {stack:3d}  SMEM {sh:6d}  LOCAL {loc:3d}  {short}")


if __name__ == "__main__":
    main()
