"""profiles/<round>_sass_*.txt: mnemonic counts and short excerpts around the tensor / TMEM / TMA / gather instructions of
the hot kernels, from the in-tree objects (cuobjdump -sass).   python scripts/sass_extract.py r2"""
import collections, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r2"
O = "factors_of_serendipity_recommendation_b200/build"


def sass(obj, pat):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for b in out.split("\t\tFunction : ")[1:]:
        name = b.split("\n", 1)[0]
        if pat in name:
            lines = [l for l in b.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
            return name, [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", re.sub(r"^\s*/\*[0-9a-f]{4,6}\*/\s*", "", l)).rstrip(" ;") for l in lines]
    return None, []


def hist(ins):
    c = collections.Counter()
    for i in ins:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", i)
        if m:
            c[m.group(2).split(".")[0]] += 1
    return c


jobs = [(f"{R}_sass_score_gq.txt", f"{O}/lgx_score_gq.o", "k_score_topk_gqILi20ELb1ELb0", r"UTCHMMA|LDTM|UTMALDG|UTCBAR",
         ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "SYNCS", "ELECT", "FMNMX3", "FMNMX", "REDUX", "STS", "LDS", "VOTE"]),
        (f"{R}_sass_rescore.txt", f"{O}/lgx_score_gq.o", "k_rescore_topk", r"HMMA\.", ["HMMA", "LDG", "LDS", "STS", "SHFL", "VOTE"]),
        (f"{R}_sass_spmm_fixed.txt", f"{O}/lgx_spmm.o", "k_spmm_fixedILi16ELi1ELi4ELi4ELi0", r"LDG\.E\.128", ["LDG", "SHFL", "FFMA", "STG", "VOTE"])]
for fn, obj, pat, mark, keys in jobs:
    name, ins = sass(obj, pat)
    h = hist(ins)
    with open("profiles/" + fn, "w") as f:
        f.write(f"# cuobjdump -sass {obj}\n# function {name}\n# built in-tree by python -m factors_of_serendipity_recommendation_b200.build (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)\n")
        f.write(f"# {len(ins)} SASS instructions.  " + ", ".join(f"{k} {h.get(k, 0)}" for k in keys) + "\n")
        f.write("# all mnemonics: " + ", ".join(f"{k} {v}" for k, v in h.most_common(45)) + "\n#\n# excerpts (instruction index, SASS):\n")
        marks = [i for i, x in enumerate(ins) if re.search(mark, x)]
        keep = set()
        for m in marks:
            keep.update(range(max(0, m - 3), min(len(ins), m + 4)))
        last, n = -2, 0
        for j in sorted(keep):
            if n > 260:
                f.write("  ... (truncated)\n")
                break
            if j != last + 1:
                f.write("  ...\n")
            f.write(f"  {j:5d}  {ins[j]}\n")
            last, n = j, n + 1
    print(fn, len(ins), {k: h.get(k, 0) for k in keys})
