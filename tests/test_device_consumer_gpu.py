"""The no-readback path (second half; VERDICT r01 missing item 2): the reference's stated
main problem is that every mesh is downloaded and re-uploaded for rendering (README.md:32-36).
lvn_meshgen_generate_batch_device leaves the exported arenas in HBM and hands out device pointers; a
graphics-API interop consumer would map its vertex / index buffers (cudaGraphicsResourceGetMappedPointer)
and copy device-to-device.  There is no GL / Vulkan in this image, so the stand-in consumer is another
CUDA library in the same process -- torch -- that wraps the arenas IN PLACE through
__cuda_array_interface__ (no copy through the host), packs every chunk's slices into its own "vertex
buffer" / "index buffer" tensors with device-to-device copies, rebases the indices the way a renderer's
draw call would (baseVertex), and only then is compared with what the host path delivers."""
import numpy as np
import pytest

import leven_b200.workloads as W

pytestmark = pytest.mark.gpu


class DeviceArray:
    """zero-copy view of a device allocation for any __cuda_array_interface__ consumer"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def test_device_arenas_consumed_in_place(lc):
    torch = pytest.importorskip("torch")
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        ms = W.ring_chunks()[192:320]                       # 128 chunks through the surface layers
        rc, res, view = ctx.generateBatchDevice(ms)
        assert rc == 0 and view.totalVertices > 0
        nV, nT = int(res["numVertices"].sum()), int(res["numTriangles"].sum())
        # the consumer's own buffers (a renderer's VBO / IBO), filled device-to-device, chunk after chunk
        vbo = torch.empty((nV, 48), dtype=torch.uint8, device="cuda")
        ibo = torch.empty((nT, 3), dtype=torch.int32, device="cuda")
        extent_v = int((res["vertexOffset"] + res["numVertices"]).max())
        extent_t = int((res["triangleOffset"] + res["numTriangles"]).max())
        dv = torch.as_tensor(DeviceArray(view.vertices, extent_v * 48), device="cuda").view(-1, 48)
        dt = torch.as_tensor(DeviceArray(view.triangles, extent_t * 12), device="cuda").view(torch.int32).view(-1, 3)
        assert dv.data_ptr() == int(view.vertices)          # in place: no staging copy was made
        v0 = t0 = 0
        draws = []
        for r in res:
            n, m = int(r["numVertices"]), int(r["numTriangles"])
            if m == 0:
                continue
            vbo[v0:v0 + n] = dv[r["vertexOffset"]:r["vertexOffset"] + n]
            ibo[t0:t0 + m] = dt[r["triangleOffset"]:r["triangleOffset"] + m] + v0      # baseVertex
            draws.append((v0, n, t0, m))
            v0 += n; t0 += m
        assert v0 == nV and t0 == nT
        # something a renderer-side pass would compute from the buffers, still on the device
        xyz = vbo.view(torch.float32).view(-1, 12)[:, :3]
        lo, hi = xyz.min(0).values.cpu().numpy(), xyz.max(0).values.cpu().numpy()
        assert int(ibo.max()) < nV and int(ibo.min()) >= 0
        # now the host path, for comparison only
        V = np.zeros(nV + 1, lc.MeshVertex); T = np.zeros(nT + 1, lc.MeshTriangle); S = np.zeros(int(view.totalSeamNodes) + 1, lc.SeamNodeInfo)
        rc, hres = ctx.generateBatch(ms, V, T, S)
        assert rc == 0
        hv = np.concatenate([V[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]] for r in hres if r["numTriangles"] > 0])
        assert vbo.cpu().numpy().tobytes() == hv.tobytes()
        base, parts = 0, []
        for r in hres:
            if r["numTriangles"] > 0:
                parts.append(T["indices_"][r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]] + base)
                base += r["numVertices"]
        assert np.array_equal(ibo.cpu().numpy(), np.concatenate(parts))
        assert np.all(lo >= ms[:, :3].min(0) - 4) and np.all(hi <= (ms[:, :3] + 256).max(0) + 4)
        assert len(draws) == int((res["numTriangles"] > 0).sum())
    finally:
        ctx.destroy()


def test_async_device_batch_counts_first(lc):
    """lvn_meshgen_generate_batch_device_async returns with the counts and offsets while the meshes are still being
    made; after lvn_meshgen_wait the arenas hold, chunk by chunk, the bytes of the synchronous call -- also when a
    consumer's work is queued behind the batch in stream order instead of waiting on the host"""
    torch = pytest.importorskip("torch")
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        ms = W.ring_chunks()
        stream = torch.cuda.current_stream()
        ctx.setStream(stream.cuda_stream)

        def arenas(res, view):
            ev = int((res["vertexOffset"] + res["numVertices"]).max()); et = int((res["triangleOffset"] + res["numTriangles"]).max())
            es = int((res["seamOffset"] + res["numSeamNodes"]).max())
            return (torch.as_tensor(DeviceArray(view.vertices, ev * 48), device="cuda").view(-1, 48),
                    torch.as_tensor(DeviceArray(view.triangles, et * 12), device="cuda").view(-1, 12),
                    torch.as_tensor(DeviceArray(view.seamNodes, max(es, 1) * 48), device="cuda").view(-1, 48))

        def per_chunk(res, dv, dt, ds):
            out = []
            for r in res:
                out.append((dv[r["vertexOffset"]:r["vertexOffset"] + r["numVertices"]].cpu().numpy().tobytes() if r["numTriangles"] else b"",
                            dt[r["triangleOffset"]:r["triangleOffset"] + r["numTriangles"]].cpu().numpy().tobytes(),
                            ds[r["seamOffset"]:r["seamOffset"] + r["numSeamNodes"]].cpu().numpy().tobytes()))
            return out

        rc, res0, view0 = ctx.generateBatchDevice(ms)
        assert rc == 0
        want = per_chunk(res0, *arenas(res0, view0))
        for rep in range(3):
            rc, res, view = ctx.generateBatchDeviceAsync(ms)
            assert rc == 0
            for k in ("numEdges", "numVertices", "numTriangles", "numSeamNodes"):
                assert (res[k] == res0[k]).all(), k
            assert view.totalVertices == view0.totalVertices and view.totalTriangles == view0.totalTriangles
            dv, dt, ds = arenas(res, view)
            if rep == 1:
                # a consumer in stream order: its copy is queued on the context's stream, no host wait in between
                snap = dv.clone()
                assert ctx.wait() == 0
                assert torch.equal(snap, dv)
            else:
                assert ctx.wait() == 0
            assert per_chunk(res, dv, dt, ds) == want
        # the next call of any kind is ordered behind an unfinished one: back-to-back counts-first calls on other
        # chunk lists, no wait in between, and the last one's arenas hold what a synchronous call delivers
        sub = ms[192:320]
        rc, res_s, view_s = ctx.generateBatchDevice(sub)
        assert rc == 0
        want_sub = per_chunk(res_s, *arenas(res_s, view_s))
        for _ in range(3):
            assert ctx.generateBatchDeviceAsync(ms)[0] == 0
            assert ctx.generateBatchDeviceAsync(ms[:64])[0] == 0
            rc, res2, view2 = ctx.generateBatchDeviceAsync(sub)
            assert rc == 0
            assert ctx.wait() == 0
            assert per_chunk(res2, *arenas(res2, view2)) == want_sub
        rc, res3, view3 = ctx.generateBatchDeviceAsync(ms)
        assert rc == 0
        rc, res4, view4 = ctx.generateBatchDevice(ms[:64])
        assert rc == 0 and (res4["numVertices"] == res0["numVertices"][:64]).all()
    finally:
        ctx.destroy()


def test_async_host_batch_counts_first(lc):
    """lvn_meshgen_generate_batch_async: results on return, host arenas after lvn_meshgen_wait, byte for byte what
    the synchronous call delivers"""
    torch = pytest.importorskip("torch")
    ctx = lc.Compute_MeshGenContext.create(64)
    try:
        ms = W.ring_chunks()
        rc, res0, view = ctx.generateBatchDevice(ms)
        assert rc == 0
        nV, nT, nS = int(view.totalVertices) + 64, int(view.totalTriangles) + 64, int(view.totalSeamNodes) + 64
        pin = [torch.zeros(n * sz, dtype=torch.uint8, pin_memory=True) for n, sz in ((nV, 48), (nT, 12), (nS, 48), (nV, 48), (nT, 12), (nS, 48))]
        V0, T0, S0 = pin[0].numpy().view(lc.MeshVertex), pin[1].numpy().view(lc.MeshTriangle), pin[2].numpy().view(lc.SeamNodeInfo)
        V1, T1, S1 = pin[3].numpy().view(lc.MeshVertex), pin[4].numpy().view(lc.MeshTriangle), pin[5].numpy().view(lc.SeamNodeInfo)
        rc, r0 = ctx.generateBatch(ms, V0, T0, S0)
        assert rc == 0
        for rep in range(2):
            rc, r1 = ctx.generateBatchAsync(ms, V1, T1, S1)
            assert rc == 0
            for k in ("numEdges", "numVertices", "numTriangles", "numSeamNodes"):
                assert (r1[k] == r0[k]).all(), k
            assert ctx.wait() == 0
            for a, b in zip(r0, r1):
                assert V0[a["vertexOffset"]:a["vertexOffset"] + a["numVertices"]].tobytes() == V1[b["vertexOffset"]:b["vertexOffset"] + b["numVertices"]].tobytes()
                assert T0[a["triangleOffset"]:a["triangleOffset"] + a["numTriangles"]].tobytes() == T1[b["triangleOffset"]:b["triangleOffset"] + b["numTriangles"]].tobytes()
                assert S0[a["seamOffset"]:a["seamOffset"] + a["numSeamNodes"]].tobytes() == S1[b["seamOffset"]:b["seamOffset"] + b["numSeamNodes"]].tobytes()
            V1[:] = 0; T1[:] = 0; S1[:] = 0
        # too small an arena is reported like by the synchronous call
        rc, r2 = ctx.generateBatchAsync(ms, V1[:1000], T1, S1)
        assert rc == lc.LVN_ERR_CAPACITY and (r2["numVertices"] == r0["numVertices"]).all()
        # right behind an edit (hash tables of the edited fields still unvalidated) the call completes the batch
        # before it returns, like the synchronous one: same meshes
        op = lc.CSGOperationInfo.make(*W.csg_script()[0])
        lo, hi = lc.CalcCSGOperationBounds(op)
        touched = W.touched_chunks(ms, lo, hi)
        assert len(touched) > 0
        assert ctx.applyCSGOperationsBatch([op], touched) == 0
        rc, ra = ctx.generateBatchAsync(touched, V1, T1, S1)
        assert rc == 0 and ctx.wait() == 0
        rc, rb = ctx.generateBatch(touched, V0, T0, S0)
        assert rc == 0 and ra["numVertices"].sum() > 0
        for a, b in zip(rb, ra):
            assert a["numVertices"] == b["numVertices"] and a["numTriangles"] == b["numTriangles"]
            assert V0[a["vertexOffset"]:a["vertexOffset"] + a["numVertices"]].tobytes() == V1[b["vertexOffset"]:b["vertexOffset"] + b["numVertices"]].tobytes()
            assert T0[a["triangleOffset"]:a["triangleOffset"] + a["numTriangles"]].tobytes() == T1[b["triangleOffset"]:b["triangleOffset"] + b["numTriangles"]].tobytes()
    finally:
        ctx.destroy()
