#!/usr/bin/env python
"""Write the CUDA source the engine would compile for a request (pm_jit_source) and, with --sass, compile it with
nvcc here (no GPU needed) and report the instruction count of the kernel.
usage: jit_dump.py [--sass] [-k 2ids] PATTERN [PATTERN2]      (default: the bench.py request)"""
import ctypes, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import patmatchdocker_b200 as pm

def source(pats, kopt):
    lib = pm.load()
    arr = (ctypes.c_char_p * len(pats))(*[p.encode() for p in pats])
    need = lib.pm_jit_source(len(pats), arr, kopt.encode(), None, 0)
    if need < 0:
        raise RuntimeError(lib.pm_last_error().decode())
    buf = ctypes.create_string_buffer(need)
    lib.pm_jit_source(len(pats), arr, kopt.encode(), buf, need)
    return buf.value.decode()

if __name__ == "__main__":
    args = sys.argv[1:]
    sass = "--sass" in args
    args = [a for a in args if a != "--sass"]
    kopt = None
    if "-k" in args:
        i = args.index("-k"); kopt = args[i + 1]; del args[i:i + 2]
    if not args:
        import bench
        args, k2 = bench.patterns()
        kopt = kopt or k2
    src = source(args, kopt or "2ids")
    out = "/tmp/jit/k.cu"
    os.makedirs("/tmp/jit", exist_ok=True)
    open(out, "w").write(src)
    print("source:", out, len(src), "bytes")
    if sass:
        subprocess.run(["nvcc", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-Xptxas", "-v", "-o", "/tmp/jit/k.cubin", out], check=True)
        txt = subprocess.run(["cuobjdump", "-sass", "/tmp/jit/k.cubin"], capture_output=True, text=True).stdout
        ops = collections.Counter()
        fn = None
        for line in txt.splitlines():
            if "Function :" in line:
                fn = line.split(":")[1].strip()
            parts = line.split()
            if len(parts) > 2 and parts[0].startswith("/*") and fn and not parts[1].startswith("/*"):
                op = parts[1] if not parts[1].startswith("@") else parts[2]
                ops[(fn, op.split(".")[0].rstrip(";"))] += 1
        for fn in sorted(set(f for f, _ in ops)):
            tot = sum(v for (f, _), v in ops.items() if f == fn)
            print(fn, "total", tot, sorted(((o, v) for (f, o), v in ops.items() if f == fn), key=lambda x: -x[1])[:12])
