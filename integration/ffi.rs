// ffi.rs — Rust binding of librt_b200.so (include/rt_b200.h, ABI version 3) for mbk6/CS397RayTracingSP22.
//
// Goes into the reference crate as `src/util/ffi.rs` (add `pub mod ffi;` to src/util.rs, see reference.patch).
// It replaces nothing by itself: `lower.rs` uses it to hand the scene that `run()` builds (tracing.rs:354-543) to the
// CUDA back end, and `Scene::render_to_image` (tracing.rs:221-263) calls `rt_render` instead of the rayon loop.
//
// NOT COMPILED in the container this repository is built in (no cargo / rustc there, crates not vendored).  The same
// symbols, with the same struct layouts, are exercised by the ctypes binding cs397raytracingsp22_b200/_ffi.py
// (every GPU test) and by include/rt_scene_api.hpp (tests/test_cpp_api.py); tests/test_integration_sources.py checks
// that every struct field and every function declared here matches the C header.
#![allow(non_camel_case_types, dead_code)]
use std::collections::HashMap;
use std::os::raw::{c_char, c_int, c_void};
use std::sync::Arc;

use super::materials::Material;
use super::texture::Texture;

pub const RT_B200_ABI_VERSION: c_int = 3;

#[repr(C)]
pub struct rt_scene { _private: [u8; 0] }

// ---- materials (materials.rs:12-166 as a tagged union)
pub const RT_MAT_LAMBERTIAN: u32 = 0;
pub const RT_MAT_METAL: u32 = 1;
pub const RT_MAT_DIELECTRIC: u32 = 2;
pub const RT_MAT_PARAMETERIZED: u32 = 3;
pub const RT_MAT_ISOTROPIC: u32 = 4;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_material_desc {
    pub tag: u32,
    pub albedo: [f32; 3],
    pub emission: [f32; 3],
    pub roughness: f32,
    pub metallic: f32,
    pub ior: f32,
}

// ---- camera: field-for-field mirror of `Camera` (tracing.rs:138-155)
pub const RT_PROJ_ORTHOGRAPHIC: u32 = 0;
pub const RT_PROJ_PERSPECTIVE: u32 = 1;
pub const RT_SHADE_PHONG: u32 = 0;
pub const RT_SHADE_PATHTRACE: u32 = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rt_camera {
    pub eyepoint: [f32; 3],
    pub view_dir: [f32; 3],
    pub up: [f32; 3],
    pub projection_mode: u32,
    pub shading_mode: u32,
    pub path_depth: u32,
    pub path_samples: u32,
    pub screen_width: u32,
    pub screen_height: u32,
    pub focal_length: f32,
    pub focus_dist: f32,
    pub lens_radius: f32,
    pub aa_sample_count: u32,
    pub max_trace_dist: f32,
    pub gamma: f32,
}

// ---- render options
pub const RT_SHARD_ALL: u32 = 0;
pub const RT_SHARD_SAMPLES: u32 = 1;
pub const RT_SHARD_TILES: u32 = 2;
pub const RT_OPT_COUNTERS: u32 = 1;
pub const RT_OPT_NO_EVENTS: u32 = 2;
pub const RT_ENGINE_AUTO: u32 = 0;
pub const RT_ENGINE_WAVEFRONT: u32 = 1;
pub const RT_ENGINE_MEGAKERNEL: u32 = 2;
pub const RT_RAYSORT_AUTO: u32 = 0;
pub const RT_RAYSORT_OFF: u32 = 1;
pub const RT_RAYSORT_ON: u32 = 2;
pub const RT_ORDER_AUTO: u32 = 0;
pub const RT_ORDER_PIXEL_MAJOR: u32 = 1;
pub const RT_ORDER_SAMPLE_MAJOR: u32 = 2;
pub const RT_ORDER_GROUPED: u32 = 3;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_render_opts {
    pub seed: u64,
    pub shard_mode: u32,
    pub shard_rank: u32,
    pub shard_count: u32,
    pub tile_size: u32,
    pub sample_begin: u32,
    pub sample_end: u32,
    pub wavefront: u32,
    pub flags: u32,
    pub point_light_pos: [f32; 3],
    pub ambient: [f32; 3],
    pub engine: u32,
    pub ray_sort: u32,
    pub work_order: u32,
    pub blocks_per_sm: u32,
    pub reserved: [u32; 4],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rt_stats {
    pub samples: u64,
    pub rays: u64,
    pub iterations: u64,
    pub kernel_launches: u64,
    pub extend_launches: u64,
    pub shade_launches: u64,
    pub nodes_visited: u64,
    pub tlas_nodes_visited: u64,
    pub tris_tested: u64,
    pub instances_entered: u64,
    pub prims_tested: u64,
    pub mesh_hits: u64,
    pub texel_taps: u64,
    pub extend_texel_taps: u64,
    pub material_fetches: u64,
    pub warp_node_slots: u64,
    pub ms_total: f64,
    pub ms_extend: f64,
    pub ms_shade: f64,
    pub ms_resolve: f64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub engine: u64,
}

extern "C" {
    pub fn rt_abi_version() -> c_int;
    pub fn rt_last_error() -> *const c_char;
    pub fn rt_device_count() -> c_int;
    pub fn rt_scene_create(out: *mut *mut rt_scene) -> c_int;
    pub fn rt_scene_destroy(s: *mut rt_scene);
    pub fn rt_add_texture(s: *mut rt_scene, rgb8: *const u8, width: u32, height: u32) -> c_int;
    pub fn rt_add_material(s: *mut rt_scene, desc: *const rt_material_desc) -> c_int;
    pub fn rt_add_mesh(s: *mut rt_scene, pos: *const f32, nrm: *const f32, uv: *const f32, nverts: u32,
                       idx: *const u32, ntris: u32) -> c_int;
    pub fn rt_add_instance(s: *mut rt_scene, mesh: c_int, xform: *const f32, inv_xform: *const f32,
                           material: c_int, tex: *const c_int) -> c_int;
    pub fn rt_add_sphere(s: *mut rt_scene, center: *const f32, radius: f32, material: c_int) -> c_int;
    pub fn rt_add_triangle(s: *mut rt_scene, a: *const f32, b: *const f32, c: *const f32, material: c_int) -> c_int;
    pub fn rt_add_plane(s: *mut rt_scene, point: *const f32, normal: *const f32, material: c_int) -> c_int;
    pub fn rt_add_volume_sphere(s: *mut rt_scene, center: *const f32, radius: f32, density: f32,
                                phase_material: c_int) -> c_int;
    pub fn rt_add_volume_mesh(s: *mut rt_scene, mesh: c_int, xform: *const f32, inv_xform: *const f32, density: f32,
                              phase_material: c_int) -> c_int;
    pub fn rt_commit(s: *mut rt_scene, device: c_int) -> c_int;
    pub fn rt_render(s: *mut rt_scene, cam: *const rt_camera, opts: *const rt_render_opts,
                     out_linear_rgb: *mut f32, out_rgb8: *mut u8, stats: *mut rt_stats) -> c_int;
    // device-resident variants for one-process-per-GPU sharding (CUDA device pointers / cudaStream_t)
    pub fn rt_render_accum(s: *mut rt_scene, cam: *const rt_camera, opts: *const rt_render_opts,
                           d_accum: *mut c_void, stream: *mut c_void, stats: *mut rt_stats) -> c_int;
    pub fn rt_resolve(s: *mut rt_scene, cam: *const rt_camera, d_accum: *const c_void, total_spp: u32,
                      d_out_linear_rgb: *mut f32, d_out_rgb8: *mut u8, stream: *mut c_void) -> c_int;
    pub fn rt_accum_bytes(width: u32, height: u32) -> usize;
}

/// The reference's own error policy is to panic (geometry.rs:149-151,168; tracing.rs:546); so does the shim.
pub fn check(rc: c_int) -> c_int {
    if rc < 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(rt_last_error()) }.to_string_lossy().into_owned();
        panic!("librt_b200 error {}: {}", rc, msg);
    }
    rc
}

/// Owns the library's scene handle while `Scene::render_to_image` lowers `Scene.objects` into it, and hands out ids.
/// `Arc`-shared materials and meshes are registered once (keyed by the address the `Arc` points to), so the 256
/// instances of BASELINE config 5 share two BLASes exactly as they share two `Arc<Mesh>` in Rust.
pub struct SceneBuilder {
    pub s: *mut rt_scene,
    materials: HashMap<usize, i32>,
    textures: HashMap<usize, i32>,
    meshes: HashMap<usize, i32>,
}

impl SceneBuilder {
    pub fn new() -> SceneBuilder {
        assert_eq!(unsafe { rt_abi_version() }, RT_B200_ABI_VERSION, "librt_b200.so and ffi.rs disagree on the ABI version");
        let mut s: *mut rt_scene = std::ptr::null_mut();
        check(unsafe { rt_scene_create(&mut s) });
        SceneBuilder { s, materials: HashMap::new(), textures: HashMap::new(), meshes: HashMap::new() }
    }

    /// `Arc<dyn Material>` -> material id (materials.rs:12-15 gains `describe()`, see lower.rs).
    pub fn material(&mut self, m: &Arc<dyn Material + Send + Sync>) -> i32 {
        let key = Arc::as_ptr(m) as *const () as usize;
        if let Some(&id) = self.materials.get(&key) {
            return id;
        }
        let desc = m.describe();
        let id = check(unsafe { rt_add_material(self.s, &desc) });
        self.materials.insert(key, id);
        id
    }

    /// `Texture` (texture.rs:11-14) -> texture id.  The decoded image crosses as RGB8, row 0 = top, which is what
    /// `Texture::sample` reads through `get_pixel(x, y).to_rgb()` (texture.rs:30).
    pub fn texture(&mut self, t: &Texture) -> i32 {
        let key = t as *const Texture as usize;
        if let Some(&id) = self.textures.get(&key) {
            return id;
        }
        let rgb = t.img.to_rgb8();
        let id = check(unsafe { rt_add_texture(self.s, rgb.as_raw().as_ptr(), rgb.width(), rgb.height()) });
        self.textures.insert(key, id);
        id
    }

    /// `Arc<tobj::Mesh>` (geometry.rs:128,157) -> mesh id; the library builds the BLAS once per mesh.
    /// tobj leaves `normals` / `texcoords` empty when the OBJ has none: zeros are passed, which is what
    /// `IndexedTriangle::intersect_ray` would index out of bounds on (geometry.rs:350-357) - the reference only
    /// loads meshes that have both.
    pub fn mesh(&mut self, m: &Arc<tobj::Mesh>) -> i32 {
        let key = Arc::as_ptr(m) as usize;
        if let Some(&id) = self.meshes.get(&key) {
            return id;
        }
        let nverts = m.positions.len() / 3;
        let zeros3;
        let zeros2;
        let nrm: &[f32] = if m.normals.len() == 3 * nverts { &m.normals } else { zeros3 = vec![0.0f32; 3 * nverts]; &zeros3 };
        let uv: &[f32] = if m.texcoords.len() == 2 * nverts { &m.texcoords } else { zeros2 = vec![0.0f32; 2 * nverts]; &zeros2 };
        let id = check(unsafe {
            rt_add_mesh(self.s, m.positions.as_ptr(), nrm.as_ptr(), uv.as_ptr(), nverts as u32, m.indices.as_ptr(),
                        (m.indices.len() / 3) as u32)
        });
        self.meshes.insert(key, id);
        id
    }
}

impl Drop for SceneBuilder {
    fn drop(&mut self) {
        unsafe { rt_scene_destroy(self.s) };
    }
}
