"""Per-kernel SASS mnemonic counts (cuobjdump -sass of the built objects): how many global / shared accesses are 128-bit wide, and
whether TMA (UTMALDG) and mbarrier (SYNCS) instructions are present. Usage: python scripts/sass_counts.py > profiles/rNN_sass_counts.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = ["dev_recon_kernels.cu.o", "dev_entropy_kernels.cu.o"]
want = ["k_reconstruct_dct8", "k_render_wide_tma", "k_render_wideI", "k_reconstructEPK", "k_renderILi1ELi1ELb1", "k_lf_groupILb1", "k_ac_vardct_multiILb1"]
print("| kernel | SASS instr | global/generic loads LDG+LD (of which .128) | stores STG+ST (.128) | LDS (.128 / .64) | STS (.128) | UTMALDG | SYNCS (mbarrier) | SHFL |\n|---|---|---|---|---|---|---|---|---|")
for o in objs:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "pdn-jpegxl_b200", "build", o)], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    for p in parts[1:]:
        name = p.split("\n", 1)[0].strip()
        if not any(w in name for w in want):
            continue
        ins = [l for l in p.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        c = lambda pat: sum(1 for l in ins if re.search(pat, l))
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().replace("jxlgpu::", "")
        print("| `%s` | %d | %d (%d) | %d (%d) | %d (%d / %d) | %d (%d) | %d | %d | %d |" % (dem[:70], len(ins), c(r"\bLDG?\.E"), c(r"\bLDG?\.E.*\.128"), c(r"\bSTG?\.E"), c(r"\bSTG?\.E.*\.128"),
              c(r"\bLDS"), c(r"\bLDS\.128"), c(r"\bLDS\.64"), c(r"\bSTS"), c(r"\bSTS\.128"), c(r"UTMALDG"), c(r"SYNCS"), c(r"SHFL")))
