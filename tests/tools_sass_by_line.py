"""Static SASS instruction counts per source line of one kernel (diagnostic, no GPU): which source regions the machine
code of a kernel comes from -- the first thing to look at before spending GPU time on an instruction-issue-bound kernel.

    python tests/tools_sass_by_line.py lep_encode_kernel [--top 40]

Counts are STATIC (one per SASS instruction); loops are not weighted.  -lineinfo attributes an instruction to the
innermost inlined source line."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def by_line(kernel_substr, defs=()):
    with tempfile.TemporaryDirectory() as td:
        cubin = os.path.join(td, "k.cubin")
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                               "-cubin", "-o", cubin] + list(defs) + [os.path.join(ROOT, "lepton_b200", "csrc", "lep_capi.cu")])
        text = subprocess.run(["/usr/local/cuda/bin/nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    counts = collections.Counter()
    inside, cur = False, ("?", 0)
    for line in text.splitlines():
        if line.startswith("//---------------------"):
            inside = ".text." in line and kernel_substr in line
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            counts[cur] += 1
    return counts


if __name__ == "__main__":
    args = sys.argv[1:]
    top = 40
    if "--top" in args:
        i = args.index("--top"); top = int(args[i + 1]); del args[i:i + 2]
    kernel = args[0]
    c = by_line(kernel, args[1:])
    total = sum(c.values())
    print("%s: %d SASS instructions" % (kernel, total))
    files = collections.Counter()
    for (f, _), n in c.items():
        files[f] += n
    for f, n in files.most_common():
        print("  %-24s %6d  %5.1f %%" % (f, n, 100.0 * n / total))
    print("top source lines:")
    for (f, l), n in c.most_common(top):
        print("  %-24s line %4d  %5d" % (f, l, n))
