"""Edge cases through the C ABI: empty and minimal inputs, zero-length runs, handles used
before anything was pushed, malformed descriptions. The reference's own tests exercise
few of these (its templates simply loop zero times); the bar here is "same values as the
oracle where there is something to compute, a clean status otherwise"."""
import ctypes as C

import numpy as np
import pytest

import wayverb_b200 as wvb
from wayverb_b200 import _lib, scene
from oracle import rto, wgo

pytestmark = pytest.mark.gpu


def to_mesh(om):
    return wvb.Mesh(om.dims, om.nodes, om.coeffs, om.b1, om.b2, om.b3)


@pytest.mark.parametrize("dims", [(5, 5, 5), (6, 5, 7), (5, 9, 5)])
def test_smallest_boxes_and_zero_length_runs(dims):
    # cuboid_inside pads by two layers: a 5^3 mesh has exactly one air node
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.3)])
    src = om.index(dims[0] // 2, dims[1] // 2, dims[2] // 2)
    with wvb.Waveguide(to_mesh(om)) as g:
        assert g.step(0) == 0 and np.count_nonzero(g.field()) == 0
        done, out, flag = g.run_device(src, np.zeros(0), [src])
        assert done == 0 and flag == 0 and out.shape == (0, 1)
        done, out, flag = g.run_device(src, np.array([1.0, 0.0, 0.0, 0.0, 0.0]), [])  # no receivers
        assert done == 5 and flag == 0
        sim = wgo.Sim(om)
        sim.run(src, np.array([1.0, 0.0, 0.0, 0.0, 0.0]), [src])
        assert np.array_equal(g.field(), sim.field())
        # a node index outside the mesh is a caller error (not "owned by another slab"): the
        # call fails with WVB_ERR_INVALID and the field is untouched
        for bad_call in (lambda: g.write(om.num_nodes + 10, 3.0), lambda: g.read(om.num_nodes + 10),
                         lambda: g.run_device(om.num_nodes, np.ones(2), [src]),
                         lambda: g.run_device(src, np.ones(2), [om.num_nodes + 3])):
            with pytest.raises(_lib.WvbError) as e:
                bad_call()
            assert "outside" in str(e.value)
        assert np.array_equal(g.field(), sim.field())


def test_mesh_without_any_boundary_or_air_node():
    dims = (8, 7, 6)
    nodes = np.zeros(dims[0] * dims[1] * dims[2], wgo.NODE_DT)  # every node id_none
    m = wvb.Mesh(dims, nodes, [wgo.to_flat(0.1)], np.zeros((0, 1), np.uint32), np.zeros((0, 2), np.uint32),
                 np.zeros((0, 3), np.uint32))
    with wvb.Waveguide(m) as g:
        g.write(m.index(3, 3, 3), 1.0)
        assert g.step(3) == 0
        assert np.count_nonzero(g.field()) == 0  # `default: return 0` (program.cpp:485)


def test_malformed_descriptions_are_rejected_with_a_message():
    dims = (8, 7, 6)
    om = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.1)])
    bad = om.nodes.copy()
    bt = bad["boundary_type"]
    first_boundary = int(np.flatnonzero((bt != 0) & ((bt & 1) == 0) & ((bt & 128) == 0))[0])
    bad["boundary_index"][first_boundary] = 10 ** 6  # points past boundary_index_array_1
    with pytest.raises(_lib.WvbError) as e:
        wvb.Waveguide(wvb.Mesh(om.dims, bad, om.coeffs, om.b1, om.b2, om.b3))
    assert e.value.status == _lib.WVB_ERR_INVALID and str(e.value)
    b1 = om.b1.copy()
    b1[0] = 99  # a coefficient set that does not exist
    with pytest.raises(_lib.WvbError):
        wvb.Waveguide(wvb.Mesh(om.dims, om.nodes, om.coeffs, b1, om.b2, om.b3))


def test_ray_handle_with_nothing_to_do():
    sc = scene.box_scene((4.0, 3.0, 6.0), subdiv=1, side=4, surfaces=[scene.make_surface(0.1, 0.1)])
    src, rcv = (1.0, 1.0, 1.0), (3.0, 2.0, 4.0)
    with wvb.RayTracer(sc) as g:
        refl, dropped, _ = g.trace(np.zeros((0, 3), np.float32), src, rcv, depth=5, n_bins=50)
        assert refl is None and dropped == 0 and np.count_nonzero(g.histogram()) == 0
        refl, _, _ = g.trace(rto.directions(1, 7), src, rcv, depth=0, n_bins=50, keep_steps=0)
        assert np.count_nonzero(g.histogram()) == 0  # depth 0: no reflection, nothing deposited
        with wvb.ImageSource(g, src, rcv, max_elements=16) as s:
            got, stats, _ = s.results()  # nothing pushed: the direct impulse alone
            want, _ = rto.image_source(rto.Scene(sc), np.zeros((1, 0), np.uint32), src, rcv)
            assert got.size == want.size == 1 and np.array_equal(got.view(np.uint8), want.view(np.uint8))
            assert stats.tolist() == [0, 0, 0, 0]
            s.push_elements(np.full((3, 4), 0xFFFFFFFF, np.uint32))  # rays that never hit anything
            got, stats, _ = s.results()
            assert got.size == 1 and stats[0] == 0
        # source == receiver: no direct impulse (get_direct.h:22-24)
        with wvb.ImageSource(g, src, src, max_elements=10) as s:
            got, _, _ = s.results()
            assert got.size == 0


def test_scene_descriptions_are_validated():
    sc = scene.box_scene((4.0, 3.0, 6.0), subdiv=1, side=4, surfaces=[scene.make_surface(0.1, 0.1)])
    sc.triangles = sc.triangles.copy()
    sc.triangles["v2"][3] = 10 ** 6
    with pytest.raises(_lib.WvbError) as e:
        wvb.RayTracer(sc)
    assert e.value.status == _lib.WVB_ERR_INVALID
    assert _lib.lib().wvb_rt_create(None, C.byref(C.c_void_p())) == _lib.WVB_ERR_INVALID
