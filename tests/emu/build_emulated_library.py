#!/usr/bin/env python
"""Builds tests/emu/libcracks_b200_emu.so: the product library's OWN sources (cracks_b200/csrc/*.cu, *.cuh)
compiled for the CPU against tests/emu/cuda_shim_full/cuda_runtime.h (TEST INFRASTRUCTURE, see that header).

The sources are copied to tests/emu/_gen/ with three mechanical rewrites, nothing else:
  1. `kernel<<<grid, block, smem, stream>>> (args);`  ->  `pf_emu::launch (coop, (kernel), grid, block, smem, stream, args);`
     where coop says whether the kernel uses barriers / shared memory / shuffles (COOPERATIVE below);
  2. `extern __shared__` -> `extern thread_local` (the dynamic shared-memory array is one buffer per OS thread,
     i.e. per running block, emu_runtime.cc);
  3. the block reductions of pf_vector.cuh go through shared memory instead of warp shuffles (speed);
  4. the tuning variants (PF_TUNING_VARIANTS, incl. the TMA kernel pf_apply3d_v3.cuh with its inline PTX) are
     not compiled, as in the product build; the reduction grid is shrunk (RED_BLOCKS x RED_THREADS = 3 x 32) because every CUDA
     thread of a cooperative kernel is an OS thread here.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "cracks_b200", "csrc")
GEN = os.path.join(HERE, "_gen")

# kernels that synchronise within a block (barriers, shared-memory staging, warp shuffles)
COOPERATIVE = ("k_apply3d", "k_diag_v6", "k_residual3d", "k_multi_dot", "k_multi_axpy_dot", "k_reduce_partials", "k_residual_finish", "k_absdiff_max",
               "k_one_minus_phi_max", "k_functionals_generic", "k_cod_generic", "k_load_top_2d")


def rewrite_launches(text):
    out, pos = [], 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            out.append(text[pos:])
            break
        # kernel expression: identifier with optional balanced template arguments, ending at i
        j = i
        if text[j - 1] == ">":
            depth = 0
            while True:
                j -= 1
                if text[j] == ">":
                    depth += 1
                elif text[j] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        kernel = text[j:i]
        k = text.index(">>>", i)
        cfg = text[i + 3:k]
        p = text.index("(", k)
        assert text[k + 3:p].strip() == "", (kernel, text[k:p + 1])
        depth, q = 0, p
        while True:
            if text[q] == "(":
                depth += 1
            elif text[q] == ")":
                depth -= 1
                if depth == 0:
                    break
            q += 1
        args = text[p + 1:q]
        ncfg = len(split_top(cfg))
        assert ncfg == 4, (kernel, cfg)
        coop = "true" if kernel.strip().startswith(COOPERATIVE) else "false"
        out.append(text[pos:j])
        out.append(f"pf_emu::launch ({coop}, ({kernel.strip()}), {cfg}, {args})")
        pos = q + 1
    return "".join(out)


def split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def generate():
    os.makedirs(GEN, exist_ok=True)
    for name in sorted(os.listdir(SRC)):
        if not name.endswith((".cu", ".cuh")) or name in ("pf_apply3d_v3.cuh", "pf_apply3d_v5.cuh"):
            continue
        text = open(os.path.join(SRC, name)).read()
        text = text.replace("extern __shared__", "extern thread_local")
        if name == "pf_vector.cuh":
            # the warp-shuffle stage of the block reductions costs ten block barriers per call when every
            # CUDA thread is an OS thread: here the reduction goes through shared memory instead (two barriers)
            for op, init, comb in (("sum", "0", "s += slots[i]"), ("max", "slots[0]", "s = fmax (s, slots[i])")):
                a = text.index(f"block_reduce_{op} (double v, double *sh")
                b = text.index("\n}\n", a) + 3
                body = (f"block_reduce_{op} (double v, double *sh)\n{{\n  static thread_local double slots[1024];\n  (void) sh;\n"
                        f"  slots[threadIdx.x] = v;\n  __syncthreads ();\n  double s = {init};\n  if (threadIdx.x == 0)\n"
                        f"    for (unsigned i = {'0' if op == 'sum' else '1'}; i < blockDim.x; ++i)\n      {comb};\n"
                        f"  __syncthreads ();\n  return s;\n}}\n")
                text = text[:a] + body + text[b:]
            text = re.sub(r"constexpr int RED_BLOCKS = \d+;", "constexpr int RED_BLOCKS = 3;", text)
            text = re.sub(r"constexpr int RED_THREADS = \d+;", "constexpr int RED_THREADS = 32;", text)
        if name == "pf_api.cu":
            # the tuning variants (incl. the TMA kernel: inline PTX, CUtensorMap) sit behind PF_TUNING_VARIANTS,
            # which this build does not define
            text = text.replace('#include "../../include/cracks_b200.h"', '#include "../../../include/cracks_b200.h"')
            text = rewrite_launches(text)
        assert "<<<" not in text, name
        open(os.path.join(GEN, name), "w").write(text)


def build():
    generate()
    so = os.path.join(HERE, "libcracks_b200_emu.so")
    cmd = ["g++", "-O1", "-g", "-fno-extern-tls-init", "-std=c++20", "-fPIC", "-shared", "-pthread", "-x", "c++", "-I", os.path.join(HERE, "cuda_shim_full"),
           "-o", so, os.path.join(GEN, "pf_api.cu"), os.path.join(HERE, "emu_runtime.cc"), "-ldl"]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build())
