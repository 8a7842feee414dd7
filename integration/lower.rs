// lower.rs — how each concrete type of the reference's scene vocabulary crosses the C ABI.
//
// Goes into the reference crate as `src/util/lower.rs` (add `pub mod lower;` to src/util.rs, see reference.patch).
// Two small traits carry the lowering; reference.patch makes them supertraits of the reference's own traits
// (`Intersectable: Lower`, tracing.rs:42-47; `Material: Describe`, materials.rs:12-15), so `Scene.objects`
// (tracing.rs:215) and every struct literal of `run()` (tracing.rs:374-543) stay exactly as they are.
//
// NOT COMPILED in the container this repository is built in (no Rust toolchain there); see ffi.rs.
use cgmath::Matrix4;
use image::RgbImage;

use super::ffi::*;
use super::geometry::{ConvexVolume, Plane, Sphere, StaticMesh, Triangle};
use super::materials::{Dielectric, Isotropic, Lambertian, Metal, ParameterizedMaterial};
use super::tracing::{CameraProjectionMode, Scene, ShadingMode, Vec3};

fn v3(v: &Vec3) -> [f32; 3] {
    [v.x, v.y, v.z]
}
fn m16(m: &Matrix4<f32>) -> [f32; 16] {
    let a: &[f32; 16] = m.as_ref(); // cgmath matrices are column-major, like the ABI
    *a
}

/// Supertrait of `Material`: the material as the tagged union of include/rt_b200.h.
pub trait Describe {
    fn describe(&self) -> rt_material_desc;
}
impl Describe for Lambertian {
    // materials.rs:20-23
    fn describe(&self) -> rt_material_desc {
        rt_material_desc { tag: RT_MAT_LAMBERTIAN, albedo: v3(&self.albedo), emission: v3(&self.emission), ..Default::default() }
    }
}
impl Describe for Metal {
    // materials.rs:51-55
    fn describe(&self) -> rt_material_desc {
        rt_material_desc { tag: RT_MAT_METAL, albedo: v3(&self.albedo), emission: v3(&self.emission), roughness: self.roughness,
                           ..Default::default() }
    }
}
impl Describe for Dielectric {
    // materials.rs:74-76 (no albedo, no emission: attenuation is 1, materials.rs:97)
    fn describe(&self) -> rt_material_desc {
        rt_material_desc { tag: RT_MAT_DIELECTRIC, ior: self.idx_of_refraction, ..Default::default() }
    }
}
impl Describe for ParameterizedMaterial {
    // materials.rs:107-112
    fn describe(&self) -> rt_material_desc {
        rt_material_desc { tag: RT_MAT_PARAMETERIZED, albedo: v3(&self.albedo), emission: v3(&self.emission),
                           roughness: self.roughness, metallic: self.metallic, ..Default::default() }
    }
}
impl Describe for Isotropic {
    // materials.rs:152-157
    fn describe(&self) -> rt_material_desc {
        rt_material_desc { tag: RT_MAT_ISOTROPIC, albedo: v3(&self.albedo), emission: v3(&self.emission), ..Default::default() }
    }
}

/// Supertrait of `Intersectable`: append this object to the scene being lowered; returns its object index.
/// Objects must be lowered in `Scene.objects` order: the index is the tie-break of the reference's linear scan
/// (tracing.rs:330-341, strict `<`: the first object wins).
pub trait Lower {
    fn lower(&self, b: &mut SceneBuilder) -> i32;
    /// What a `ConvexVolume` needs to know about its boundary (geometry.rs:496): only a Sphere or a StaticMesh can
    /// produce the exit hit of geometry.rs:508-509.
    fn as_sphere(&self) -> Option<(Vec3, f32)> {
        None
    }
    fn as_static_mesh(&self) -> Option<&StaticMesh> {
        None
    }
}
impl Lower for Sphere {
    // geometry.rs:389-393
    fn lower(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material(&self.material);
        check(unsafe { rt_add_sphere(b.s, v3(&self.center).as_ptr(), self.radius, m) })
    }
    fn as_sphere(&self) -> Option<(Vec3, f32)> {
        Some((self.center, self.radius))
    }
}
impl Lower for Triangle {
    // geometry.rs:424-429
    fn lower(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material(&self.material);
        check(unsafe { rt_add_triangle(b.s, v3(&self.a).as_ptr(), v3(&self.b).as_ptr(), v3(&self.c).as_ptr(), m) })
    }
}
impl Lower for Plane {
    // geometry.rs:468-472
    fn lower(&self, b: &mut SceneBuilder) -> i32 {
        let m = b.material(&self.material);
        check(unsafe { rt_add_plane(b.s, v3(&self.point).as_ptr(), v3(&self.normal).as_ptr(), m) })
    }
}
impl Lower for ConvexVolume {
    // geometry.rs:495-500.  The boundary's own material is ignored, as in geometry.rs:505-510.
    fn lower(&self, b: &mut SceneBuilder) -> i32 {
        let phase = b.material(&self.phase_function);
        if let Some((c, r)) = self.boundary.as_sphere() {
            return check(unsafe { rt_add_volume_sphere(b.s, v3(&c).as_ptr(), r, self.density, phase) });
        }
        let sm = self.boundary.as_static_mesh().expect("ConvexVolume: the boundary must be a Sphere or a StaticMesh");
        let mesh = b.mesh(&sm.mesh);
        // inv_xform = NULL: the library inverts the affine transform itself (cofactors, like cgmath's inverse_transform,
        // geometry.rs:168); passing sm.inv_transform works too (its last row is accepted within a few ulps of 0 0 0 1)
        check(unsafe { rt_add_volume_mesh(b.s, mesh, m16(&sm.transform).as_ptr(), std::ptr::null(), self.density, phase) })
    }
}
impl Lower for StaticMesh {
    // geometry.rs:127-134: mesh (shared BLAS), optional fixed material, five optional textures, transform
    fn lower(&self, b: &mut SceneBuilder) -> i32 {
        let mesh = b.mesh(&self.mesh);
        let mut tex = [-1i32; 5];
        for (k, t) in self.textures.iter().enumerate() {
            if let Some(t) = t {
                tex[k] = b.texture(t);
            }
        }
        let mat = match &self.material {
            Some(m) => b.material(m),
            None => -1, // textures drive a ParameterizedMaterial per hit (geometry.rs:253-271)
        };
        check(unsafe {
            rt_add_instance(b.s, mesh, m16(&self.transform).as_ptr(), m16(&self.inv_transform).as_ptr(), mat, tex.as_ptr())
        })
    }
    fn as_static_mesh(&self) -> Option<&StaticMesh> {
        Some(self)
    }
}

impl Scene {
    /// The body of `Scene::render_to_image` (tracing.rs:221-263) on the CUDA back end: lower, commit, render, wrap the
    /// bytes.  reference.patch makes `render_to_image` call this; everything else in tracing.rs stays.
    pub fn render_to_image_b200(&self, device: i32, seed: u64) -> RgbImage {
        let mut b = SceneBuilder::new();
        for object in self.objects.iter() {
            object.lower(&mut b);
        }
        check(unsafe { rt_commit(b.s, device) });
        let c = &self.camera;
        let cam = rt_camera {
            eyepoint: v3(&c.eyepoint),
            view_dir: v3(&c.view_dir),
            up: v3(&c.up),
            projection_mode: match c.projection_mode {
                CameraProjectionMode::Orthographic => RT_PROJ_ORTHOGRAPHIC,
                CameraProjectionMode::Perspective => RT_PROJ_PERSPECTIVE,
            },
            shading_mode: match c.shading_mode {
                ShadingMode::Phong => RT_SHADE_PHONG,
                ShadingMode::PathTrace => RT_SHADE_PATHTRACE,
            },
            path_depth: c.path_depth,
            path_samples: c.path_samples,
            screen_width: c.screen_width,
            screen_height: c.screen_height,
            focal_length: c.focal_length,
            focus_dist: c.focus_dist,
            lens_radius: c.lens_radius,
            aa_sample_count: c.aa_sample_count,
            max_trace_dist: c.max_trace_dist,
            gamma: c.gamma,
        };
        let opts = rt_render_opts { seed, point_light_pos: v3(&self.point_light_pos), ambient: v3(&self.ambient),
                                    ..Default::default() };
        let mut bytes = vec![0u8; (c.screen_width as usize) * (c.screen_height as usize) * 3];
        let mut stats = rt_stats::default();
        check(unsafe { rt_render(b.s, &cam, &opts, std::ptr::null_mut(), bytes.as_mut_ptr(), &mut stats) });
        println!("librt_b200: {} samples, {} rays, {:.1} ms on the device", stats.samples, stats.rays, stats.ms_total);
        RgbImage::from_raw(c.screen_width, c.screen_height, bytes).expect("image buffer size")
    } // `b` drops here: rt_scene_destroy
}
