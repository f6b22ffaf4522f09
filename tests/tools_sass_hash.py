"""Per-kernel SASS fingerprints of lepton_b200/csrc/lep_capi.cu (diagnostic, no GPU): instruction count and an md5 of the
instruction text of every kernel, so that "the default kernels are still the ones that were validated on the GPU" is a
command, not a claim.

    python tests/tools_sass_hash.py                  # print the fingerprints of the working tree
    python tests/tools_sass_hash.py --check profiles/r01_sass_validated.txt      # compare with a recorded set
    python tests/tools_sass_hash.py -DLEPB200_STREAM_HINTS=1                     # any build-time option
"""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fingerprints(defs=()):
    with tempfile.TemporaryDirectory() as td:
        cubin = os.path.join(td, "k.cubin")
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                               "-cubin", "-o", cubin] + list(defs) + [os.path.join(ROOT, "lepton_b200", "csrc", "lep_capi.cu")])
        sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    cur, out = None, {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+", "", m.group(1))          # anonymous-namespace hash varies per build
            out[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if m and cur:
            out[cur].append(m.group(1))
    return {k: (len(v), hashlib.md5("\n".join(v).encode()).hexdigest()) for k, v in out.items()}


if __name__ == "__main__":
    args = sys.argv[1:]
    check = None
    if "--check" in args:
        i = args.index("--check")
        check = args[i + 1]
        del args[i:i + 2]
    fp = fingerprints(args)
    if check is None:
        for k in sorted(fp):
            print(k, fp[k][0], fp[k][1])
        sys.exit(0)
    want = {}
    for line in open(check):
        if line.strip() and not line.startswith("#"):
            k, n, h = line.split()
            want[k] = (int(n), h)
    bad = 0
    for k in sorted(want):
        same = fp.get(k) == want[k]
        bad += not same
        print("SAME" if same else "DIFF", k[:70], want[k][0], fp.get(k, ("-",))[0])
    for k in sorted(set(fp) - set(want)):
        print("NEW ", k[:70], fp[k][0])
    sys.exit(1 if bad else 0)
