#!/usr/bin/env python
"""Make one of the reference's OpenCL C sources palatable to g++ (with clc.hpp), without changing
what it computes.  Usage: translate.py <in> <out.inc> [--drop-include <header> ...]

Only syntax that C++ cannot parse is rewritten; every expression, table and constant stays the
reference's own text (a #line directive points compiler diagnostics back at it):
  (float4)(a, b, c, d)   ->  float4(a, b, c, d)        OpenCL vector literal -> constructor call
  #include "cl/x.glsl"   ->  #include "x.glsl.inc"     the translated copy next to this one
  #pragma OPENCL / unroll ->  dropped
For the reference's C++ (leven/src/octree.cpp, the seam octree) nothing is rewritten at all;
--drop-include removes #include lines of headers whose own includes cannot be satisfied here
(window system, renderer, thread pool: render.h, volume.h, ...) -- the two functions and one
constant octree.cpp takes from them are supplied by oracle/ref_shim/ref_octree.cpp.
--lines A-B keeps only that (1-based, inclusive) slice of a file, verbatim, behind a #line directive:
clipmap.cpp's mesh-data functions (GenerateMeshDataForNode .. ConstructCollisionNodeData) are compiled
that way against include/leven_compute.hpp, without the renderer / physics / thread code around them.
The output is a build intermediate under oracle/_ref/ (git-ignored, deleted after the build).
"""
import os
import re
import sys


def translate(text, src_path, drop=(), ranges=()):
    if ranges:
        lines = text.split("\n")
        parts = []
        for a, b in ranges:
            parts.append('#line %d "%s"' % (a, src_path))
            parts.extend(lines[a - 1:b])
        return "\n".join(parts) + "\n"
    for h in drop:
        text = re.sub(r'^[ \t]*#include\s+"%s".*$' % re.escape(h), "// (include of %s dropped)" % h, text, flags=re.M)
    if src_path.endswith((".cpp", ".h")):
        # MSVC's __m128 is a union with a float array member, gcc's is a vector type: x.m128_f32[i] -> x[i]
        text = re.sub(r"\.m128_f32\s*\[", "[", text)
        return '#line 1 "%s"\n%s\n' % (src_path, text)
    text = re.sub(r"\(\s*(float|int|uint)([234])\s*\)\s*\(", r"\1\2(", text)
    text = re.sub(r'#include\s+"cl/([A-Za-z0-9_.]+)"', r'#include "\1.inc"', text)
    text = re.sub(r"^\s*#pragma\s+(OPENCL|unroll).*$", "", text, flags=re.M)
    return '#line 1 "%s"\n%s\n' % (src_path, text)


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    drop = [sys.argv[i + 1] for i, a in enumerate(sys.argv) if a == "--drop-include"]
    # --lines A-B (repeatable): only that slice of a C++ source, verbatim (clipmap.cpp's mesh-data functions)
    ranges = [tuple(int(v) for v in sys.argv[i + 1].split("-")) for i, a in enumerate(sys.argv) if a == "--lines"]
    with open(src, encoding="utf-8", errors="replace") as f:
        out = translate(f.read(), os.path.abspath(src), drop, ranges)
    with open(dst, "w") as f:
        f.write(out)
