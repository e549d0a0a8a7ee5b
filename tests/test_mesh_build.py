"""Mesh construction: compute_mesh's node part on the device
against the oracle's restatement of the same reference kernels.

CPU part: the oracle's voxel_inside / closest-triangle restatements against
analytic answers for a shoebox. GPU part: wvb_mesh_create == oracle, node for node."""
import numpy as np
import pytest

import wayverb_b200 as wvb
from wayverb_b200 import scene
from oracle import rto, wgo

BOX = (4.0, 3.0, 6.0)


def descriptor(spacing=0.25, pad=2):
    # like compute_adjusted_boundary (boundary_adjust.cpp:8-22): the mesh AABB encloses the
    # room with spare layers
    mc = np.array([-pad * spacing + 0.01] * 3, np.float32)
    dims = tuple(int(np.ceil((b - float(mc[0])) / spacing)) + pad for b in BOX)
    return mc, dims, np.float32(spacing)


def two_surface_room():
    return scene.box_scene(BOX, subdiv=2, side=8, per_wall_surfaces=True,
                           surfaces=[scene.make_surface(0.1, 0.1), scene.make_surface(0.2, 0.1),
                                     scene.make_surface(0.3, 0.1)])


def test_oracle_inside_matches_box_analytically():
    sc = two_surface_room()
    o = rto.Scene(sc)
    mc, dims, sp = descriptor()
    ins = o.nodes_inside(mc, dims, sp)
    z, y, x = np.indices(ins.shape)
    px, py, pz = (mc[0] + x.astype(np.float32) * sp, mc[1] + y.astype(np.float32) * sp,
                  mc[2] + z.astype(np.float32) * sp)
    want = (px > 0) & (px < BOX[0]) & (py > 0) & (py < BOX[1]) & (pz > 0) & (pz < BOX[2])
    assert want.sum() > 1000
    assert np.array_equal(ins, want)


def test_oracle_closest_surface_picks_the_nearest_wall():
    sc = two_surface_room()
    o = rto.Scene(sc)
    pts = np.array([[2.0, 1.5, 0.1], [2.0, 1.5, 5.9], [0.1, 1.5, 3.0], [3.9, 1.5, 3.0], [2.0, 0.1, 3.0],
                    [2.0, 2.9, 3.0]], np.float32)
    surf, tri = o.closest_surface(pts)
    # box_scene: walls z=0, z=max, x=0, x=max, y=0, y=max get surfaces 0,1,2,0,1,2
    assert surf.tolist() == [0, 1, 2, 0, 1, 2]


@pytest.mark.gpu
def test_device_mesh_from_scene_matches_oracle():
    sc = two_surface_room()
    o = rto.Scene(sc)
    mc, dims, sp = descriptor()
    coeffs = [wgo.to_flat(0.1), wgo.to_flat(0.2), wgo.to_flat(0.3)]
    with wvb.RayTracer(sc) as g:
        m, ins = wvb.build_mesh(dims, mc, sp, coeffs, scene=g, return_inside=True)
    want_ins = o.nodes_inside(mc, dims, sp)
    assert np.array_equal(ins.astype(bool), want_ins)
    z, y, x = np.indices(want_ins.shape)
    pts = np.stack([mc[0] + x.astype(np.float32) * sp, mc[1] + y.astype(np.float32) * sp,
                    mc[2] + z.astype(np.float32) * sp], -1).reshape(-1, 3)
    surf, _ = o.closest_surface(pts)
    om = wgo.mesh_from_inside(want_ins, coeffs, surf)
    assert np.array_equal(m.nodes["boundary_type"], om.nodes["boundary_type"])
    assert np.array_equal(m.nodes["boundary_index"], om.nodes["boundary_index"])
    assert np.array_equal(m.b[0], om.b1) and np.array_equal(m.b[1], om.b2) and np.array_equal(m.b[2], om.b3)
    assert len(set(m.b[0].ravel().tolist())) == 3
    # and the mesh it built runs: same field as the oracle's mesh
    src = m.index(dims[0] // 2, dims[1] // 2, dims[2] // 2)
    sim = wgo.Sim(om)
    sim.write(src, 1.0)
    with wvb.Waveguide(m) as w:
        w.write(src, 1.0)
        assert sim.step(40) == 0 and w.step(40) == 0
        assert np.array_equal(w.field(), sim.field())


@pytest.mark.gpu
def test_device_mesh_from_mask_with_reentrant_nodes():
    ins = np.zeros((12, 14, 14), bool)
    ins[2:10, 2:12, 2:7] = True
    ins[2:10, 2:7, 2:12] = True
    zz, yy, xx = np.indices(ins.shape)
    surf = ((xx > 6).astype(np.uint32) + (yy > 6).astype(np.uint32)).ravel()
    coeffs = [wgo.to_flat(0.1), wgo.to_flat(0.2), wgo.to_flat(0.3)]
    m = wvb.build_mesh((14, 14, 12), (0, 0, 0), 0.1, coeffs, inside=ins, surface_1d=surf)
    om = wgo.mesh_from_inside(ins, coeffs, surf)
    assert (om.nodes["boundary_type"] == wgo.ID_REENTRANT).any()
    assert np.array_equal(m.nodes, om.nodes)
    assert np.array_equal(m.b[0], om.b1) and np.array_equal(m.b[1], om.b2) and np.array_equal(m.b[2], om.b3)
