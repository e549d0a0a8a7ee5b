"""Pins the oracle's mesh-construction restatement to the REFERENCE'S OWN kernel source:
set_node_inside / set_node_boundary_type (src/waveguide/src/mesh_setup_program.cpp:110-172) and
boundary_coefficient_finder_1d/2d/3d (boundary_coefficient_program.cpp:310-484), compiled from
/root/reference into oracle/_ref by oracle/ref_recipe/build.py. Node for node, index for index."""
import numpy as np
import pytest

from oracle import refk, rto, wgo
from wayverb_b200 import scene

pytestmark = pytest.mark.skipif(not refk.available(), reason="no /root/reference and no prebuilt oracle/_ref")

BOX = (4.0, 3.0, 6.0)


def descriptor(spacing=0.25, pad=2):
    mc = np.array([-pad * spacing + 0.01] * 3, np.float32)
    dims = tuple(int(np.ceil((b - float(mc[0])) / spacing)) + pad for b in BOX)
    return mc, dims, np.float32(spacing)


def room():
    return scene.box_scene(BOX, subdiv=2, side=8, per_wall_surfaces=True,
                           surfaces=[scene.make_surface(0.1, 0.1), scene.make_surface(0.2, 0.1),
                                     scene.make_surface(0.3, 0.1)])


def number(nodes):
    """boundary_index numbering on the host (boundary_coefficient_finder.cpp:12-19,39-64):
    running count per popcount class over non-inside, non-reentrant nodes, in node order."""
    out = nodes.copy()
    bt = out["boundary_type"]
    pc = np.array([bin(int(v) & 0xFFFFFFFF).count("1") for v in bt])
    boundary = (bt != wgo.ID_NONE) & (bt != wgo.ID_INSIDE) & (bt != wgo.ID_REENTRANT)
    counts = []
    for k in (1, 2, 3):
        sel = boundary & (pc == k)
        out["boundary_index"][sel] = np.arange(sel.sum(), dtype=np.uint32)
        counts.append(int(sel.sum()))
    return out, counts


def test_scene_classification_and_coefficient_indices():
    sc = room()
    mc, dims, sp = descriptor()
    o = rto.Scene(sc)
    ins = o.nodes_inside(mc, dims, sp)
    z, y, x = np.indices(ins.shape)
    pts = np.stack([mc[0] + x.astype(np.float32) * sp, mc[1] + y.astype(np.float32) * sp,
                    mc[2] + z.astype(np.float32) * sp], -1).reshape(-1, 3)
    surf, _ = o.closest_surface(pts)
    coeffs = [wgo.to_flat(0.1), wgo.to_flat(0.2), wgo.to_flat(0.3)]
    om = wgo.mesh_from_inside(ins, coeffs, surf)

    ref_nodes = refk.classify_scene(sc, mc, dims, sp)
    assert np.array_equal(ref_nodes["boundary_type"] == wgo.ID_INSIDE, ins.ravel())
    assert np.array_equal(ref_nodes["boundary_type"], om.nodes["boundary_type"])
    numbered, (n1, n2, n3) = number(ref_nodes)
    assert np.array_equal(numbered["boundary_index"], om.nodes["boundary_index"])
    assert (n1, n2, n3) == (om.b1.shape[0], om.b2.shape[0], om.b3.shape[0])
    b1, b2, b3 = refk.coefficient_indices(sc, numbered, mc, dims, sp, n1, n2, n3)
    assert np.array_equal(b1, om.b1) and np.array_equal(b2, om.b2) and np.array_equal(b3, om.b3)
    assert len(set(b1.ravel().tolist())) == 3 and n2 > 0 and n3 == 8


def test_boundary_types_of_a_mask_with_reentrant_nodes():
    ins = np.zeros((12, 14, 14), bool)
    ins[2:10, 2:12, 2:7] = True
    ins[2:10, 2:7, 2:12] = True
    nodes = np.zeros(ins.size, refk.NODE_DT)
    nodes["boundary_type"][ins.ravel()] = wgo.ID_INSIDE
    got = refk.boundary_type_only(nodes, (14, 14, 12))
    want = wgo.classify(ins)
    assert (want["boundary_type"] == wgo.ID_REENTRANT).any()
    assert np.array_equal(got["boundary_type"], want["boundary_type"])


def test_inside_test_on_an_l_shaped_solid():
    """voxel_inside's 32-direction parity test (voxel.cpp:156-225) on a non-convex solid."""
    # an L-shaped prism out of two boxes' outer faces is awkward to triangulate by hand; use two
    # overlapping closed boxes: a point inside both crosses 2 surfaces in every direction -> the
    # reference calls it OUTSIDE (parity). The oracle must reproduce exactly that.
    a = scene.box_scene((2.0, 2.0, 2.0), subdiv=1, side=4)
    va = a.vertices[:, :3]
    vb = va * np.float32(0.5) + np.float32(0.75)
    tris = np.concatenate([a.triangles, a.triangles.copy()])
    tris[a.triangles.size:]["v0"] += va.shape[0]
    tris[a.triangles.size:]["v1"] += va.shape[0]
    tris[a.triangles.size:]["v2"] += va.shape[0]
    sc = scene.Scene(np.concatenate([va, vb]), tris, a.surfaces, side=4)
    mc, dims, sp = np.array([-0.3, -0.3, -0.3], np.float32), (14, 14, 14), np.float32(0.2)
    ins = rto.Scene(sc).nodes_inside(mc, dims, sp)
    ref_nodes = refk.classify_scene(sc, mc, dims, sp)
    assert np.array_equal(ref_nodes["boundary_type"] == wgo.ID_INSIDE, ins.ravel())
    assert ins.any() and not ins[7, 7, 7]   # the doubly-enclosed core counts as outside
