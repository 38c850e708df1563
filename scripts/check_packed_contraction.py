"""ptxas (12.9) contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 although both carry an explicit rounding
modifier (it never does that to the scalar `.rn` forms).  The two-nodes-per-thread kernels are bit-identical to the
one-node kernels only if the source leaves no such pair, i.e. every fusable multiply-add is spelled `vfma`.

This script compiles every collision operator for float2 into a stand-alone kernel and compares, per kernel, the
number of packed fused multiply-adds in the PTX (fma.rn.f32x2) with the genuine FFMA2 of the SASS.  ptxas rewrites
FMUL2 -> FFMA2 (x * y + -0) and FADD2 -> FFMA2 (x * 1 + y) freely, which keeps the bits; those carry an RZ or a
literal 1 operand and are not counted.  More genuine FFMA2 than fma.rn.f32x2 means a contraction.  Exit status 1 if
any kernel was contracted.  (No GPU needed.)

    python scripts/check_packed_contraction.py
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CASES = [(s, c) for s in ("D2Q9", "D3Q19", "D3Q27")
         for c in ("LBM_OP_BGK", "LBM_OP_TRT", "LBM_OP_KBC", "LBM_OP_REGULARIZED", "LBM_OP_SMAGORINSKY", "FORCED")
         if not (s == "D3Q19" and c == "LBM_OP_KBC")]

SRC_HEAD = '#include "%s/lettuce_b200/csrc/lbm_core.cuh"\nusing namespace lbm;\n' % ROOT
KERNEL = """
extern "C" __global__ void k_%(name)s(float2 *f, float a, float b, ForceArgs<float> fa) {
    float2 v[%(S)s::Q];
    for (int q = 0; q < %(S)s::Q; ++q) v[q] = f[q * 1000 + threadIdx.x];
    %(call)s
    for (int q = 0; q < %(S)s::Q; ++q) f[q * 1000 + threadIdx.x] = v[q];
}
"""


def main():
    src = SRC_HEAD
    names = []
    for s, c in CASES:
        name = f"{s}_{c}"
        names.append(name)
        call = (f"collide_bgk_forced<{s}, float2>(v, a, fa);" if c == "FORCED"
                else f"Collide<{s}, float2, {c}>::apply(v, a, b);")
        src += KERNEL % dict(name=name, S=s, call=call)
    with tempfile.TemporaryDirectory() as tmp:
        cu = os.path.join(tmp, "t.cu")
        open(cu, "w").write(src)
        flags = ["-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                 "-I", os.path.join(ROOT, "include")]
        subprocess.run([NVCC, *flags, "-ptx", cu, "-o", os.path.join(tmp, "t.ptx")], check=True)
        subprocess.run([NVCC, *flags, "-cubin", cu, "-o", os.path.join(tmp, "t.cubin")], check=True)
        ptx = open(os.path.join(tmp, "t.ptx")).read()
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, "t.cubin")], capture_output=True, text=True,
                              check=True).stdout
    bad = 0
    for name in names:
        m = re.search(r"\.entry k_%s\((.*?)\n}\n" % name, ptx, re.S)
        body = m.group(1)
        n_ptx = len(re.findall(r"\bfma\.rn\.f32x2", body))
        n_all = len(re.findall(r"\b(?:add|sub|mul|fma)\.rn\.f32x2", body))
        m = re.search(r"Function : k_%s\n(.*?)(?=Function :|\Z)" % name, sass, re.S)
        n_sass = n_pseudo = 0
        for ins in re.findall(r"\bFFMA2 ([^;]*);", m.group(1)):
            ops = [o.strip().lstrip("-") for o in ins.split(",")][1:]
            # a * b + c: a rewritten add has a literal-1 factor, a rewritten mul a zero addend
            if len(ops) == 3 and (ops[0] == "1" or ops[1] == "1" or ops[2] in ("RZ", "RZ.F32")):
                n_pseudo += 1
            else:
                n_sass += 1
        n_sass_all = len(re.findall(r"\b(?:FADD2|FMUL2|FFMA2)\b", m.group(1)))
        flag = "" if n_ptx == n_sass else f"   <-- {n_sass - n_ptx} contraction(s)"
        bad += n_ptx != n_sass
        print(f"{name:32s} fma.rn.f32x2 {n_ptx:4d}  genuine FFMA2 {n_sass:4d}  (packed ops: ptx {n_all}, sass {n_sass_all}, "
              f"rewritten mul/add {n_pseudo}){flag}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
