"""cuobjdump -sass opcode histograms of the built objects -> profiles/<round>_sass_opcodes.txt (run here, no GPU needed).

  python tools/sass_histogram.py profiles/r2_sass_opcodes.txt
"""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tiktaalik_b200", "csrc")
OBJS = ["rk4.o", "apply.o", "assemble.o", "evolve.o", "gk.o", "exchange.o", "capi.o"]
HEAD = ("SASS opcode histograms of the built objects (cuobjdump -sass, static counts per function; nvcc 12.9, -gencode arch=compute_100a,code=sm_100a).\n"
        "DMMA.8x8x4 = mma.sync.m8n8k4.f64 (FP64 tensor cores: tcgen05 has no FP64 kind), LDGSTS = cp.async, BAR.* named barriers of the K2 pipeline, "
        "MUFU.RCP64H the seed of every IEEE division.\n")


def histogram(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    per, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            per[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            key = op if op.startswith(("BAR", "MUFU", "DMMA", "LDGSTS", "MEMBAR")) else op.split(".")[0]
            per[name][key] += 1
    return per


def main(path):
    with open(path, "w") as f:
        f.write(HEAD)
        for obj in OBJS:
            if not os.path.exists(os.path.join(CSRC, obj)):
                continue
            per = histogram(obj)
            tot = collections.Counter()
            for c in per.values():
                tot.update(c)
            f.write("\n==== %s (sm_100a)\n  all functions: %d instructions; %s\n"
                    % (obj, sum(tot.values()), ", ".join("%s %d" % kv for kv in tot.most_common(28))))
            for name, c in per.items():
                f.write("  %-110s %6d instr: %s\n" % (name[:110], sum(c.values()), ", ".join("%s %d" % kv for kv in c.most_common(10))))
                special = {k: v for k, v in c.items() if k.startswith(("BAR", "DMMA", "LDGSTS", "MUFU.RCP64H", "LDL", "STL"))}
                if special:
                    f.write("  %-110s        of which: %s\n" % ("", ", ".join("%s %d" % kv for kv in sorted(special.items()))))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt"))
