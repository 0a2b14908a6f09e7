"""A/B helper: builds libsvb200 from the csrc of a git revision into ab/<name>/libsvb200.so (git-ignored, ships to the GPU
box), so that one gpurun call can time two builds: SVB_LIB=ab/<name>/libsvb200.so python tools/prof_stage.py k4 1024 10
Usage: python tools/build_variant.py <git-rev | WORK> <name> [extra nvcc flags...]   (WORK = the working tree)"""
import concurrent.futures as cf
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sudoku-vision_b200"))
import build as B  # noqa: E402

rev, name, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
out_dir = os.path.join(ROOT, "ab", name)
os.makedirs(out_dir, exist_ok=True)
with tempfile.TemporaryDirectory() as tmp:
    if rev == "WORK":
        subprocess.check_call(f"cd {ROOT} && tar -c sudoku-vision_b200/csrc include | tar -x -C {tmp}", shell=True)
    else:
        subprocess.check_call(f"git -C {ROOT} archive {rev} sudoku-vision_b200/csrc include | tar -x -C {tmp}", shell=True)
    csrc = os.path.join(tmp, "sudoku-vision_b200", "csrc")
    srcs = [os.path.join(csrc, s) for s in B.SOURCES if os.path.exists(os.path.join(csrc, s))]

    def comp(s):
        o = s[:-3] + ".o"
        subprocess.check_call([B._nvcc()] + B.NVCC_FLAGS + extra + ["-c", "-o", o, s])
        return o

    with cf.ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(comp, srcs))
    so = os.path.join(out_dir, "libsvb200.so")
    subprocess.check_call([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", so] + objs)
    print(so)
