"""The whole chain a wayverb user drives, on the device, against the oracle chain on the
CPU (every stage already has its own parity test; this one checks that they compose):

  scene -> mesh (wvb_mesh_*, mesh.cpp:53-141)
        -> per-surface boundary filters (wvb_lrs_*, mesh.cpp:126-138 / fitted_boundary.h:79-104)
        -> waveguide run with the device-side source/receiver (wvb_wg_run, waveguide.h:36-126)
  scene -> rays (wvb_rt_*, raytracer.h:188-266)
        -> stochastic histogram + image-source impulses (wvb_is_*, reflection_processor/*)

Waveguide: bit-identical receiver trace. The designed filters enter both sides from the
library (the design is host code with its own test against the oracle, test_lrs_design.py)."""
import numpy as np
import pytest

import wayverb_b200 as wvb
from wayverb_b200 import lrs, scene
from oracle import rto, wgo

pytestmark = pytest.mark.gpu

BOX = (2.4, 1.8, 3.0)
ABSORPTION = ([0.08, 0.08, 0.2, 0.5, 0.4, 0.4, 0.36, 0.3], [0.15, 0.15, 0.11, 0.1, 0.07, 0.06, 0.06, 0.05],
              [0.02, 0.02, 0.03, 0.03, 0.03, 0.04, 0.07, 0.08])


def test_scene_to_impulse_responses_matches_oracle_chain():
    surfaces = [scene.make_surface(a, 0.05) for a in ABSORPTION]
    sc = scene.box_scene(BOX, subdiv=2, side=8, per_wall_surfaces=True, surfaces=surfaces)
    o = rto.Scene(sc)
    spacing, pad, c_sound = np.float32(0.1), 2, 340.0
    mc = np.array([-pad * float(spacing) + 0.013] * 3, np.float32)
    dims = tuple(int(np.ceil((b - float(mc[0])) / float(spacing))) + pad for b in BOX)
    fs = 1.0 / (float(spacing) / (c_sound * np.sqrt(3.0)))  # compute_sample_rate, config.cpp:19-21
    # mesh.cpp:126-138: one impedance filter per surface
    coeffs = [lrs.to_impedance_coefficients(lrs.compute_reflectance_filter_coefficients(a, fs)) for a in ABSORPTION]
    assert all(lrs.is_stable(c["a"]) for c in coeffs)

    with wvb.RayTracer(sc) as g:
        mesh, inside = wvb.build_mesh(dims, mc, spacing, coeffs, scene=g, return_inside=True)
        want_inside = o.nodes_inside(mc, dims, spacing)
        assert np.array_equal(inside.astype(bool), want_inside)
        z, y, x = np.indices(want_inside.shape)
        pts = np.stack([mc[0] + x.astype(np.float32) * spacing, mc[1] + y.astype(np.float32) * spacing,
                        mc[2] + z.astype(np.float32) * spacing], -1).reshape(-1, 3)
        surf, _ = o.closest_surface(pts)
        om = wgo.mesh_from_inside(want_inside, coeffs, surf)
        assert np.array_equal(mesh.nodes["boundary_type"], om.nodes["boundary_type"])

        # waveguide: calibrated impulse at the source node, pressure at the receiver node
        loc = lambda p: [int(np.round((np.float32(v) - mc[0]) / spacing)) for v in p]  # noqa: E731
        source, receiver = (0.7, 0.9, 1.1), (1.6, 0.8, 2.0)
        src, rcv = mesh.index(*loc(source)), mesh.index(*loc(receiver))
        steps = 300
        sig = np.zeros(steps)
        sig[0] = np.float32(np.sqrt(400.0 / (4 * np.pi)) / (0.3405 * float(spacing)))
        with wvb.Waveguide(mesh) as w:
            done, got, flag = w.run_device(src, sig, [rcv])
        wdone, want, wflag = wgo.Sim(om).run(src, sig, [rcv])
        assert done == wdone == steps and flag == wflag == 0
        assert np.abs(want).max() > 0 and np.array_equal(got, want)
        assert np.abs(want[200:]).max() < np.abs(want[:100]).max()  # the fitted walls absorb

        # rays + image sources in the same scene
        n, order = 20000, 4
        depth = wvb.reflection_depth(min(min(a) for a in ABSORPTION))
        depth = min(depth, 40)
        dirs = rto.directions(17, n)
        want_h, want_r, _ = o.trace(dirs, source, receiver, depth, seed=17, keep_steps=order, n_bins=800,
                                    specular_from_step=order + 1)
        want_i, _ = rto.image_source(o, rto.path_elements(want_r, order), source, receiver)
        with wvb.ImageSource(g, source, receiver, max_elements=n * order) as s:
            s.trace(dirs, depth=depth, order=order, seed=17, n_bins=800, specular_from_step=order + 1)
            got_h = g.histogram()
            got_i, stats, _ = s.results()
        assert np.abs(got_h - want_h).max() <= 1e-9 * np.abs(want_h).max()
        assert got_i.shape == want_i.shape and np.array_equal(got_i.view(np.uint8), want_i.view(np.uint8))
        assert got_i.size > 7 and stats[2] == 0
