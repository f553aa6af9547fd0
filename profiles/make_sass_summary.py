"""Count the Blackwell-specific SASS mnemonics of libcoopcap.so (tcgen05 MMA = UTCHMMA / UTCQMMA,
TMEM loads = LDTM, TMA = UTMALDG / UTMASTG / UTMAREDG, ...), whole library and per kernel.

    python profiles/make_sass_summary.py > profiles/r02_final_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "cooperativeimagecaptioning_b200", "libcoopcap.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTMAPF|SYNCS|LDGSTS|HMMA|IMMA|MUFU)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    cur, per, tot = None, collections.OrderedDict(), collections.Counter()
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for k in PAT.findall(ln):
            per[cur][k] += 1
            tot[k] += 1
    print("SASS mnemonic counts of cooperativeimagecaptioning_b200/libcoopcap.so (cuobjdump -sass, sm_100a)")
    print("whole library: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items(), key=lambda x: -x[1])))
    print("\nkernels that touch the tensor core / TMEM / TMA:")
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    for (fn, c), name in zip(per.items(), names):
        if any(k in c for k in ("UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG")):
            name = re.sub(r"\(.*", "", name)[:120]
            print(f"  {name}: " + ", ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda x: -x[1])))


if __name__ == "__main__":
    main()
