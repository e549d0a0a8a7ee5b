"""CPU-side checks of the product's host logic (no GPU compute):
the C-ABI library loads and exports everything include/wvb200.h declares, the
synthetic cuboid generator agrees with the reference's classification rules
(restated in the oracle), slab bookkeeping, and loud failure without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import wayverb_b200 as wvb
from wayverb_b200 import _lib
from oracle import wgo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        if h.endswith(".h"):
            src = open(os.path.join(ROOT, "include", h)).read()
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            names |= set(re.findall(r"\b(wvb_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    decl = declared_symbols()
    assert {"wvb_wg_create", "wvb_wg_step", "wvb_wg_run", "wvb_mesh_cuboid"} <= decl
    for name in sorted(decl):
        assert hasattr(L, name), "libwvb200.so does not export " + name
    assert L.wvb_version() == 100
    for name in (_lib.WG_SYMBOLS + _lib.RT_SYMBOLS + _lib.MESH_SYMBOLS + _lib.IS_SYMBOLS + _lib.LRS_SYMBOLS +
                 _lib.SCENE_SYMBOLS + _lib.PP_SYMBOLS):
        assert name in decl


def test_pod_layouts():
    assert _lib.NODE_DT.itemsize == 8 and _lib.COEFF_DT.itemsize == 112 and _lib.BDATA_DT.itemsize == 56
    assert C.sizeof(_lib.WgRunParams) == 72
    # wvb_is_desc: 2 x float[3], double, 2 x int32, uint64; raytracer::impulse<8> is 64 bytes
    assert C.sizeof(_lib.IsDesc) == 48 and _lib.IsDesc.acoustic_impedance.offset == 24
    assert _lib.IMPULSE_DT.itemsize == 64 and _lib.IMPULSE_DT.fields["distance"][1] == 48


def test_header_struct_sizes_with_the_c_compiler(tmp_path):
    """the ctypes mirrors against the header itself: compile a C program that prints sizeof"""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "wvb200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(wvb_is_desc), sizeof(wvb_impulse), sizeof(wvb_rt_trace_params), sizeof(wvb_wg_run_params),'
                   'sizeof(wvb_coefficients_canonical), sizeof(wvb_reflection), sizeof(wvb_condensed_node));return 0;}\n')
    exe = tmp_path / "sizes"
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-o", str(exe), str(src)],
                   check=True)  # the header is plain C
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(_lib.IsDesc), _lib.IMPULSE_DT.itemsize, C.sizeof(_lib.RtTraceParams),
                   C.sizeof(_lib.WgRunParams), _lib.COEFF_DT.itemsize, 32, _lib.NODE_DT.itemsize]


@pytest.mark.parametrize("dims", [(14, 12, 10), (5, 5, 5), (9, 6, 7), (33, 27, 22)])
def test_cuboid_generator_matches_reference_classification(dims):
    nodes, counts = wvb.waveguide.cuboid_nodes(dims)
    ref = wgo.mesh_from_inside(wgo.cuboid_inside(dims), [wgo.to_flat(0.1)])
    assert np.array_equal(nodes["boundary_type"], ref.nodes["boundary_type"])
    # the reference leaves stale indices on inside nodes only when they were
    # numbered at some point; for a box they never are, so indices match too
    assert np.array_equal(nodes["boundary_index"], ref.nodes["boundary_index"])
    assert counts == (ref.b1.shape[0], ref.b2.shape[0], ref.b3.shape[0])


def test_cuboid_slabs_concatenate_to_whole():
    dims = (12, 9, 17)
    whole, counts = wvb.waveguide.cuboid_nodes(dims)
    parts = []
    for r in range(4):
        z0, z1 = wvb.slab_range(dims[2], r, 4)
        p, c = wvb.waveguide.cuboid_nodes(dims, z0, z1 - z0)
        assert c == counts
        parts.append(p)
    assert np.array_equal(np.concatenate(parts), whole)


def test_slab_range_partitions():
    for dz in (7, 64, 513):
        for n in (1, 2, 3, 8):
            r = [wvb.slab_range(dz, k, n) for k in range(n)]
            assert r[0][0] == 0 and r[-1][1] == dz
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_cuboid_rejects_tiny():
    with pytest.raises(_lib.WvbError):
        wvb.waveguide.cuboid_nodes((4, 8, 8))


def test_no_cpu_fallback():
    """Without a usable GPU the product must fail loudly, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = wvb.cuboid_mesh((8, 8, 8), [wgo.to_flat(0.1)])
    with pytest.raises(_lib.WvbError) as e:
        wvb.Waveguide(m)
    assert e.value.status in (_lib.WVB_ERR_NO_DEVICE, _lib.WVB_ERR_CUDA)


def test_rt_host_helpers_and_validation():
    from wayverb_b200 import scene
    from oracle import rto
    # compute_optimum_reflection_number: ceil(-6 / log10(1 - a)); 132 for a = 0.1
    assert wvb.reflection_depth(0.1) == 132
    src, rcv = [1.0, 1.0, 1.0], [2.0, 3.0, 1.5]
    assert wvb.raytracer.ray_energy(1000, src, rcv, 0.1) == rto.ray_energy(1000, src, rcv, 0.1)
    sc = scene.box_scene(subdiv=1, side=4)
    bad = scene.Scene(sc.vertices[:, :3], sc.triangles, sc.surfaces, side=4)
    bad.triangles["v0"][0] = 10_000
    with pytest.raises(_lib.WvbError) as e:
        wvb.RayTracer(bad)
    assert e.value.status == _lib.WVB_ERR_INVALID


def test_create_validates_description():
    m = wvb.cuboid_mesh((8, 8, 8), [wgo.to_flat(0.1)])
    with pytest.raises(_lib.WvbError) as e:
        wvb.Waveguide(m, z_range=(4, 2))
    assert e.value.status == _lib.WVB_ERR_INVALID
    part = wvb.cuboid_mesh((8, 8, 8), [wgo.to_flat(0.1)], z0=3, nz=2)
    with pytest.raises(_lib.WvbError) as e:
        wvb.Waveguide(part, z_range=(2, 6))  # nodes do not cover the slab + ghosts
    assert e.value.status == _lib.WVB_ERR_INVALID
