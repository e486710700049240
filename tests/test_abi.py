"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol
include/leven_compute.h declares, and the POD layouts are the reference's."""
import ctypes as C
import os
import sys
import re
import subprocess
import tempfile

import numpy as np

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "leven_compute.h")


def _declared_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lvn_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built):
    import leven_b200.compute as lc
    L = lc.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in leven_compute.h but not exported"
    assert set(declared) == set(lc.ABI.keys())


def test_no_torch_or_oracle_in_product():
    """the product never routes through the oracle (or any CPU fallback)"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "leven_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle"
                assert "lvo_" not in src


def test_pod_layouts(built):
    """sizeof/offsetof from the C header via gcc == the reference PODs == the numpy dtypes"""
    import leven_b200.compute as lc
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "leven_compute.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(lvn_csg_operation_info), sizeof(lvn_seam_node_info),
         sizeof(lvn_mesh_vertex), sizeof(lvn_mesh_triangle), sizeof(lvn_aabb), sizeof(lvn_chunk_result));
  printf("%zu %zu %zu\n", offsetof(lvn_csg_operation_info, origin), offsetof(lvn_csg_operation_info, dimensions),
         offsetof(lvn_seam_node_info, position));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.c")
        open(src, "w").write(prog)
        exe = os.path.join(td, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = list(map(int, out))
    # CSGOperationInfo 48 (compute.h:16-24), SeamNodeInfo 48 (compute.h:26-31), MeshVertex 48
    # (render_types.h:24-39), MeshTriangle 12 (render_types.h:42-58), AABB 24 (aabb.h:95-96)
    assert sizes[:6] == [48, 48, 48, 12, 24, 32]
    assert sizes[6:] == [16, 32, 16]
    assert C.sizeof(lc.CSGOperationInfo) == 48 and C.sizeof(lc.AABB) == 24
    assert lc.MeshVertex.itemsize == 48 and lc.MeshTriangle.itemsize == 12 and lc.SeamNodeInfo.itemsize == 48
    assert lc.ChunkResult.itemsize == 32


def test_cpp_shim_compiles(built):
    """the header-only C++ shim (compute.h's own class on top of the C ABI) compiles as C++11"""
    prog = r'''
#include "leven_compute.hpp"
int main() {
  static_assert(sizeof(CSGOperationInfo) == 48, "CSGOperationInfo");
  static_assert(sizeof(SeamNodeInfo) == 48, "SeamNodeInfo");
  static_assert(sizeof(MeshVertex) == 48, "MeshVertex");
  static_assert(sizeof(MeshTriangle) == 12, "MeshTriangle");
  Compute_MeshGenContext* (*create)(const int) = &Compute_MeshGenContext::create;
  (void)create;
  void (*simplify)(MeshBuffer*, const lvn_shim::vec4&, const MeshSimplificationOptions&) = &ngMeshSimplifier;   // ng_mesh_simplify.h:32-36
  (void)simplify;
  static_assert(sizeof(lvn_simplify_options) == 24 && sizeof(lvn_simplify_job) == 32 && sizeof(lvn_clipmap_node) == 24, "simplify / update PODs");
  return GetCLErrorString(0) == nullptr;
}'''
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "t.cpp")
        open(src, "w").write(prog)
        subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), src])


def test_no_device_fails_loudly(built):
    """without a CUDA device the product reports LVN_ERR_NO_DEVICE; it never computes on the CPU"""
    import leven_b200.compute as lc
    import torch
    if torch.cuda.is_available():
        return
    assert lc.Compute_Initialise(1, 0, 2) == lc.LVN_ERR_NO_DEVICE
    assert lc.Compute_MeshGenContext.create(64).privateCtx_ is None
    assert lc.GetCLErrorString(lc.LVN_ERR_NO_DEVICE) == "LVN_ERR_NO_DEVICE"
    assert lc.FindNextPrime(2048) == 2053        # host-only helper works anywhere
    # the widened rows likewise: no device, no result
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import simplify_scenarios as S
    v, t = S.torus()
    rc, out, res = lc.ngMeshSimplifierBatch([(v, t, [0, 0, 0])], lc.SimplifyOptions.make())
    assert rc == lc.LVN_ERR_NO_DEVICE and len(out[0][1]) == 0
    assert lc.GenerateClipmapSeamMeshes(64, [([0, 0, 0], 256, [])])[0] == lc.LVN_ERR_NO_DEVICE
    # pass 2 of the update: the seam-update set is host logic (it is counted before the device is
    # asked for anything): a 2 x 2 x 2 block of newly active nodes invalidates all eight seams, a
    # ninth node two cells away none; the contouring itself then has no device to run on
    nodes = np.zeros(9, lc.ClipmapNode)
    nodes["min"][:8] = [[x * 256, y * 256, z * 256] for x in (0, 1) for y in (0, 1) for z in (0, 1)]
    nodes["min"][8] = [1024, 1024, 1024]
    nodes["size"] = 256
    nodes["numSeamNodes"] = 1
    everyone = np.arange(9, dtype=np.int32)
    V = np.zeros(16, lc.MeshVertex); T = np.zeros(16, lc.MeshTriangle); arena = np.zeros(16, lc.SeamNodeInfo)
    rc, upd, sres, n_all = lc.ClipmapSeamUpdateBatch(64, nodes, everyone, everyone[:8], arena, 16, V, T, 1, 2)
    assert rc == lc.LVN_ERR_NO_DEVICE and n_all == 8 and upd.tolist() == [1, 3, 5, 7]
    rc, upd, sres, n_all = lc.ClipmapSeamUpdateBatch(64, nodes, everyone, everyone[8:], arena, 16, V, T)
    assert n_all == 1 and upd.tolist() == [8]
    assert lc.ClipmapSeamUpdateBatch(64, nodes, everyone, everyone[:0], arena, 16, V, T)[::3] == (0, 0)
    assert lc.ClipmapSeamUpdateBatch(64, nodes, everyone, everyone, arena, 16, V, T, 2, 2)[0] == lc.LVN_ERR_INVALID_VALUE


def test_packed_fp32_is_not_contracted(built):
    """Arithmetic spec guard (density.cuh): ptxas contracts a packed FMUL2 feeding an FADD2 into
    FFMA2 even under --fmad=false, so every packed product is written fma(a, b, -0) with an opaque
    -0.  The library's SASS must then hold packed adds and fmas but not one FMUL2; every written
    packed add must still be there (360 per inlined pair evaluation) and at least the written
    FFMA2s (301 per evaluation: 252 in the 14 snoise2, 49 in the fractals; ptxas may
    rematerialise a frequency product under register pressure, which repeats the same rounding)."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "leven_b200", "lib", "libleven_b200.so")
    sass = subprocess.check_output([cuobjdump, "-sass", lib]).decode()
    assert "FMUL2" not in sass
    per_fn, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_fn[cur] = {"FFMA2": 0, "FADD2": 0}
        elif cur:
            for op in ("FFMA2", "FADD2"):
                if re.search(r"\b%s\b" % op, line):
                    per_fn[cur][op] += 1
    herm = [v for k, v in per_fn.items() if "k_hermite_terrain" in k]
    cols = [v for k, v in per_fn.items() if "k_columns" in k]
    assert len(herm) == 2 and len(cols) == 1   # the run-time-geometry instance and the V = 64 one
    assert cols[0]["FFMA2"] == 301 and cols[0]["FADD2"] == 360, cols[0]
    # two inlined evaluations (phase A, phase B)
    for h in herm:
        assert h["FFMA2"] >= 2 * 301 and h["FADD2"] == 2 * 360, h
    # the split kernels: one evaluation each
    for name in ("k_hermite_search", "k_hermite_normals"):
        ks = [v for k, v in per_fn.items() if name in k]
        assert len(ks) == 1 and ks[0]["FFMA2"] >= 301 and ks[0]["FADD2"] == 360, (name, ks)


def test_kernel_register_budgets(built):
    """the register counts the measured occupancies rest on: a launch-bounds edit that lets ptxas take
    254 registers for a noise kernel passes every parity test and quarters its occupancy"""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "leven_b200", "lib", "libleven_b200.so")
    out = subprocess.check_output([cuobjdump, "-res-usage", lib]).decode()
    regs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+?):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and cur:
            regs[cur] = int(m.group(1))
            cur = None
    budgets = {"k_hermite_terrainILi64": 64, "k_hermite_terrainILi0": 64, "k_leavesILi64": 72, "k_rowsILi64": 32, "k_solveE": 48,
               "k_columnsE": 40, "9k_hermiteE": 72, "k_field_densityE": 64, "k_csg_emitE": 64}
    for key, limit in budgets.items():
        hits = {k: v for k, v in regs.items() if key in k}
        assert hits, key
        for k, v in hits.items():
            assert v <= limit, (k, v, limit)
