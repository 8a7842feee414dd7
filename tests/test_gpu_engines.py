"""The two engines of the CUDA path - the wavefront pipeline and the persistent megakernel - must produce the SAME
accumulator, bit for bit: same device functions, same Philox keys, integer accumulation (include/rt_b200.h,
RT_ENGINE_*).  So every parity statement the other tests make about one engine holds for the other."""
import numpy as np
import pytest

import cs397raytracingsp22_b200 as rt
from cs397raytracingsp22_b200 import _ffi, distributed as D

pytestmark = pytest.mark.gpu
SEED = 0x5EED
WF, MK = _ffi.RT_ENGINE_WAVEFRONT, _ffi.RT_ENGINE_MEGAKERNEL


def _accum(g, cam, opts_list):
    import torch
    acc = D.new_accum(cam.screen_width, cam.screen_height, torch.device("cuda", 0))
    stats = [D.render_shard(g, cam, o, acc) for o in opts_list]
    torch.cuda.synchronize()
    return acc, stats


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
def test_megakernel_equals_wavefront_bit_for_bit(gpu, small_scenes, name):
    import torch
    sc = small_scenes(name)
    g = sc.commit(0)
    cam = sc.camera.to_c()
    wf, st_w = _accum(g, cam, [D.shard_opts(0, 1, SEED, "all", engine=WF)])
    mk, st_m = _accum(g, cam, [D.shard_opts(0, 1, SEED, "all", engine=MK)])
    assert int(wf.abs().sum()) > 0
    assert torch.equal(wf, mk), f"{name}: {int((wf != mk).sum())} accumulator words differ"
    assert st_w[0].samples == st_m[0].samples == cam.screen_width * cam.screen_height * cam.aa_sample_count
    assert st_w[0].rays == st_m[0].rays
    assert st_m[0].kernel_launches == 2 and st_m[0].iterations == 1
    # the ray sort and the wavefront width do not change the result either
    for kw in (dict(ray_sort=_ffi.RT_RAYSORT_ON), dict(ray_sort=_ffi.RT_RAYSORT_OFF), dict(wavefront=4096)):
        other, _ = _accum(g, cam, [D.shard_opts(0, 1, SEED, "all", engine=WF, **kw)])
        assert torch.equal(wf, other), (name, kw)
    # resident blocks per SM is a scheduling knob, not a result knob
    few, _ = _accum(g, cam, [D.shard_opts(0, 1, SEED, "all", engine=MK, blocks_per_sm=1)])
    assert torch.equal(wf, few)


@pytest.mark.parametrize("mode,world,kw", [("samples", 3, {}), ("tiles", 4, dict(tile=32)), ("tiles", 3, dict(tile=48))])
def test_megakernel_shards_sum_to_the_frame(gpu, small_scenes, mode, world, kw):
    import torch
    sc = small_scenes("c4")
    g = sc.commit(0)
    cam = sc.camera.to_c()
    full, _ = _accum(g, cam, [D.shard_opts(0, 1, SEED, "all", engine=WF)])
    parts, sts = _accum(g, cam, [D.shard_opts(r, world, SEED, mode, engine=MK, **kw) for r in range(world)])
    assert torch.equal(full, parts)
    assert sum(s.samples for s in sts) == cam.screen_width * cam.screen_height * cam.aa_sample_count


def test_megakernel_with_a_mesh_bounded_volume(gpu):
    """k_path<VOLMESH>: ConvexVolume whose boundary is a StaticMesh (nested boundary queries inside the traversal)."""
    import torch
    from cs397raytracingsp22_b200 import cgmath as cg, scenes
    cam = rt.Camera(eyepoint=(0.0, 1.0, 4.0), screen_width=96, screen_height=64, aa_sample_count=16, path_depth=6)
    fog = rt.ConvexVolume(
        boundary=rt.StaticMesh.load_from_file(scenes.obj_path("cube"), material=rt.Lambertian(),
                                              transform=cg.chain(cg.from_translation((0.0, 1.0, 0.0)), cg.from_angle_y(30.0))),
        phase_function=rt.Isotropic(albedo=(0.9, 0.8, 0.7)), density=1.5)
    light = rt.Lambertian(albedo=(0.0, 0.0, 0.0), emission=(4.0, 4.0, 4.0))
    sc = rt.Scene(camera=cam, objects=[fog, rt.Plane((0, 0, 0), (0, 1, 0), rt.Lambertian(albedo=(0.5, 0.5, 0.5))),
                                       rt.Sphere((0.0, 4.0, 0.0), 1.0, light)])
    g = sc.commit(0)
    c = cam.to_c()
    wf, _ = _accum(g, c, [D.shard_opts(0, 1, SEED, "all", engine=WF)])
    mk, _ = _accum(g, c, [D.shard_opts(0, 1, SEED, "all", engine=MK)])
    assert int(wf.abs().sum()) > 0 and torch.equal(wf, mk)


def test_auto_engine_and_refusals(gpu, small_scenes):
    """AUTO must pick an engine that supports the call; asking the megakernel for a wavefront-only mode is an error."""
    sc = small_scenes("c1")
    g = sc.commit(0)
    cam = sc.camera.to_c()
    o = _ffi.rt_render_opts(); o.seed = SEED
    lin_a, _, st_a = g.render(cam, o)                      # AUTO on a scene without big meshes: the megakernel
    assert st_a.kernel_launches <= 4
    o.engine = WF
    lin_w, _, st_w = g.render(cam, o)
    assert st_w.kernel_launches > st_a.kernel_launches and np.array_equal(lin_a, lin_w)
    o.engine = MK
    o.flags = _ffi.RT_OPT_COUNTERS
    with pytest.raises(_ffi.RtError) as e:
        g.render(cam, o)
    assert e.value.code == _ffi.RT_ERR_UNSUPPORTED
    o.flags = 0
    ph = sc.camera.to_c()
    ph.shading_mode = _ffi.RT_SHADE_PHONG
    with pytest.raises(_ffi.RtError) as e:
        g.render(ph, o)
    assert e.value.code == _ffi.RT_ERR_UNSUPPORTED
    br = sc.camera.to_c()
    br.path_samples = 2
    br.path_depth = 3
    with pytest.raises(_ffi.RtError) as e:
        g.render(br, o)
    assert e.value.code == _ffi.RT_ERR_UNSUPPORTED
    o.engine = _ffi.RT_ENGINE_AUTO                         # AUTO falls back to the wavefront engine for these
    g.render(ph, o)
    g.render(br, o)
    o.engine = 7
    with pytest.raises(_ffi.RtError) as e:
        g.render(cam, o)
    assert e.value.code == _ffi.RT_ERR_INVALID
    o.engine = 0
    o.reserved[2] = 1
    with pytest.raises(_ffi.RtError):
        g.render(cam, o)


def test_checkpointed_render_resumes_bit_for_bit(gpu, small_scenes, tmp_path):
    """rt_render_progressive: the host accumulator is the checkpoint.  Render sample indices [0, 5),
    write the accumulator to disk, throw the scene handle away, load the file in a fresh handle, continue with [5, 16):
    the result is the image - and the accumulator - of an uninterrupted 16-spp render, bit for bit."""
    from cs397raytracingsp22_b200 import scenes
    from conftest import SMALL
    sc = scenes.make_scene("c4", **SMALL["c4"])           # private scene objects: their handles are closed below
    cam = sc.camera.to_c()
    spp = cam.aa_sample_count
    g = sc.commit(0)
    o = _ffi.rt_render_opts(); o.seed = SEED
    lin_ref, rgb_ref, _ = g.render(cam, o)
    acc = np.zeros(cam.screen_height * cam.screen_width * 4, np.int64)
    o.sample_begin, o.sample_end = 0, 5
    lin5, _, st = g.render_progressive(cam, o, acc, 5)
    assert st.samples == cam.screen_width * cam.screen_height * 5 and np.isfinite(lin5).all()
    np.save(tmp_path / "ckpt.npy", acc)
    sc.close()                                            # "the process dies"
    sc2 = scenes.make_scene("c4", **SMALL["c4"])          # a second, independently lowered scene
    g2 = sc2.commit(0)
    acc2 = np.load(tmp_path / "ckpt.npy")
    o.sample_begin, o.sample_end = 5, spp
    lin, rgb, _ = g2.render_progressive(cam, o, acc2, spp)
    assert np.array_equal(lin, lin_ref) and np.array_equal(rgb, rgb_ref)
    # and the accumulator equals the one-shot accumulator
    one, _ = _accum(g2, cam, [D.shard_opts(0, 1, SEED, "all")])
    assert np.array_equal(one.cpu().numpy(), acc2)
    sc2.close()
