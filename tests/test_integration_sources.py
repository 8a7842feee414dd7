"""integration/ffi.rs and integration/lower.rs are the Rust shim a maintainer of the reference adds.
There is no Rust toolchain in this image, so they cannot be compiled here; what CAN be checked is that the binding
says the same thing as the C header: every struct has the same fields in the same order with types of the same size,
every extern function exists in the header with the same number of parameters, and lower.rs implements the lowering
for every concrete type of the reference's scene vocabulary."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rt_b200.h")).read(), flags=re.S)
FFI = re.sub(r"//[^\n]*", "", open(os.path.join(ROOT, "integration", "ffi.rs")).read())
LOWER = open(os.path.join(ROOT, "integration", "lower.rs")).read()

C_SIZE = {"uint64_t": 8, "uint32_t": 4, "float": 4, "double": 8}
RS_SIZE = {"u64": 8, "u32": 4, "f32": 4, "f64": 8}


def c_struct(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HDR, re.S).group(1)
    out = []
    for ty, field, arr in re.findall(r"(\w+)\s+(\w+)(?:\[(\d+)\])?;", body):
        out.append((field, C_SIZE[ty] * int(arr or 1)))
    return out


def rs_struct(name):
    body = re.search(r"pub struct %s \{(.*?)\n\}" % name, FFI, re.S).group(1)
    out = []
    for field, ty in re.findall(r"pub (\w+):\s*([^,\n]+),", body):
        m = re.fullmatch(r"\[(\w+); (\d+)\]", ty.strip())
        out.append((field, RS_SIZE[m.group(1)] * int(m.group(2)) if m else RS_SIZE[ty.strip()]))
    return out


def test_struct_layouts_match_the_header():
    for name in ("rt_material_desc", "rt_camera", "rt_render_opts", "rt_stats"):
        assert c_struct(name) == rs_struct(name), name
    # and the ctypes binding agrees with both
    import ctypes as C
    from cs397raytracingsp22_b200 import _ffi
    for name in ("rt_material_desc", "rt_camera", "rt_render_opts", "rt_stats"):
        cls = getattr(_ffi, name)
        assert [(f, C.sizeof(t)) for f, t in cls._fields_] == c_struct(name), name


def test_extern_functions_match_the_header():
    block = re.search(r'extern "C" \{(.*?)\n\}', FFI, re.S).group(1)
    fns = re.findall(r"pub fn (rt_\w+)\((.*?)\)", block, re.S)
    assert len(fns) >= 19
    for name, params in fns:
        m = re.search(r"\b%s\s*\((.*?)\);" % name, HDR, re.S)
        assert m, f"{name} is not declared in rt_b200.h"
        c_params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        rs_params = [p for p in params.split(",") if p.strip()]
        assert len(c_params) == len(rs_params), name
    consts = dict(re.findall(r"pub const (RT_\w+): (?:u32|c_int) = (\d+);", FFI))
    for k, v in consts.items():
        m = re.search(r"\b%s\s*=\s*(1u << \d+|\d+)" % k, HDR) or re.search(r"#define %s (\d+)" % k, HDR)
        assert m, k
        val = m.group(1)
        val = 1 << int(val.split("<<")[1]) if "<<" in val else int(val)
        assert val == int(v), k


def test_every_scene_type_of_the_reference_is_lowered():
    for ty in ("Sphere", "Triangle", "Plane", "ConvexVolume", "StaticMesh"):           # geometry.rs
        assert re.search(r"impl Lower for %s \{" % ty, LOWER), ty
    for ty in ("Lambertian", "Metal", "Dielectric", "ParameterizedMaterial", "Isotropic"):   # materials.rs
        assert re.search(r"impl Describe for %s \{" % ty, LOWER), ty
    for fn in ("pub fn new()", "pub fn material(", "pub fn texture(", "pub fn mesh(", "impl Drop for SceneBuilder"):
        assert fn in FFI, fn
    assert "render_to_image_b200" in LOWER and "rt_commit" in LOWER and "rt_render(" in LOWER
    assert "..." not in re.sub(r"\.\.Default::default\(\)", "", FFI + LOWER), "no elided bodies in the shipped sources"
