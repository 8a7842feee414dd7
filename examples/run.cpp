// examples/run.cpp — the C++ counterpart of the reference's `run()` (tracing.rs:354-548): build a scene with the
// mirrored API, render it on the GPU through the C ABI, write render.png.  Built by tests/test_cpp_api.py:
//   g++ -std=c++17 -Iinclude examples/run.cpp -o build/rt_run cs397raytracingsp22_b200/librt_b200.so -Wl,-rpath,...
// usage: rt_run [obj_dir [out.png|out.tga [width height spp [texture_dir]]]]
#include <cstdio>
#include <cstdlib>
#include <string>

#include "rt_scene_api.hpp"

using namespace rt;

int main(int argc, char** argv) {
  std::string obj_dir = argc > 1 ? argv[1] : "./obj";
  std::string out = argc > 2 ? argv[2] : "render.png";
  try {
    Scene scene;
    scene.camera.screen_width = argc > 3 ? (uint32_t)std::atoi(argv[3]) : 100;   // tracing.rs:366-368
    scene.camera.screen_height = argc > 4 ? (uint32_t)std::atoi(argv[4]) : 100;
    scene.camera.aa_sample_count = argc > 5 ? (uint32_t)std::atoi(argv[5]) : 100;
    auto obj = [&](auto v) { scene.objects.push_back(std::make_shared<decltype(v)>(std::move(v))); };
    // meshes of run() (tracing.rs:375-404).  The Drone_*.tga maps are missing upstream; like the reference,
    // Texture::load_from_file quietly returns None then and the Q7 defaults apply.
    std::string tex = argc > 6 ? std::string(argv[6]) + "/" : obj_dir + "/../texture/";
    obj(StaticMesh::load_from_file(obj_dir + "/drone.obj", (tex + "Drone_Albedo.tga").c_str(), (tex + "Drone_Emission.tga").c_str(),
                                   (tex + "Drone_Metallic.tga").c_str(), (tex + "Drone_Roughness.tga").c_str(),
                                   (tex + "Drone_Normal.tga").c_str(), nullptr,
                                   Matrix4::from_translation(vec3(0.0f, 1.3f, 1.7f)) * Matrix4::from_angle_y(-60.0f) *
                                       Matrix4::from_angle_x(180.0f) * Matrix4::from_scale(0.0030f)));
    obj(StaticMesh::load_from_file(obj_dir + "/cube.obj", (tex + "green.png").c_str(), nullptr, nullptr, nullptr,
                                   (tex + "normal_test.jpg").c_str(), nullptr,
                                   Matrix4::from_translation(vec3(-1.7f, 0.5f, 2.7f)) * Matrix4::from_angle_y(45.0f) * Matrix4::from_scale(0.4f)));
    obj(StaticMesh::load_from_file(obj_dir + "/sphere.obj", (tex + "magenta.jpg").c_str(), nullptr, nullptr, nullptr,
                                   (tex + "normal_test.png").c_str(), nullptr,
                                   Matrix4::from_translation(vec3(1.7f, 0.5f, 2.7f)) * Matrix4::from_angle_y(45.0f) * Matrix4::from_scale(0.6f)));
    // demo of the parameterized material (tracing.rs:406-483)
    const float ys[3] = {3.3f, 4.4f, 5.5f}, metallic[3] = {0.0f, 0.5f, 1.0f};
    const float xs[5] = {-2.6f, -1.3f, 0.0f, 1.3f, 2.6f}, rough[5] = {0.0f, 0.25f, 0.5f, 0.75f, 1.0f};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 5; ++c)
        obj(Sphere(vec3(xs[c], ys[r], 0.0f), 0.5f,
                   std::make_shared<ParameterizedMaterial>(vec3(0.01f, 0.02f, 0.5f), vec3(0, 0, 0), rough[c], metallic[r])));
    obj(Sphere(vec3(-2.3f, 2.0f, 2.0f), 0.4f, std::make_shared<Dielectric>(2.5f)));
    obj(Sphere(vec3(2.3f, 2.0f, 2.0f), 0.4f, std::make_shared<Lambertian>(vec3(0.3f, 0.3f, 0.3f), vec3(0.0f, 1.0f, 1.0f))));
    obj(ConvexVolume(std::make_shared<Sphere>(vec3(-3.0f, 1.0f, 1.0f), 1.0f, std::make_shared<Dielectric>(1.5f)),
                     std::make_shared<Isotropic>(vec3(1, 1, 1)), 0.6f));
    obj(ConvexVolume(std::make_shared<Sphere>(vec3(3.0f, 1.0f, 1.0f), 1.0f, std::make_shared<Dielectric>(1.5f)),
                     std::make_shared<Isotropic>(vec3(0, 0, 0)), 0.8f));
    obj(Plane(vec3(0, 0, 0), vec3(0, 1, 0), std::make_shared<ParameterizedMaterial>(vec3(0.33f, 0.33f, 0.33f), vec3(0, 0, 0), 0.7f, 0.3f)));
    auto light = std::make_shared<Lambertian>(vec3(0.0f, 0.6f, 0.0f), vec3(7.0f, 7.0f, 7.0f));
    obj(Triangle(vec3(-2.5f, 7.5f, -0.5f), vec3(2.5f, 7.5f, -0.5f), vec3(2.5f, 7.5f, 3.5f), light));
    obj(Triangle(vec3(-2.5f, 7.5f, -0.5f), vec3(-2.5f, 7.5f, 3.5f), vec3(2.5f, 7.5f, 3.5f), light));

    rt_stats st;
    RgbImage img = scene.render_to_image(0, &st);
    if (out.size() > 4 && out.compare(out.size() - 4, 4, ".tga") == 0) img.save_tga(out);
    else img.save_png(out);  // tracing.rs:546
    std::printf("rendered %ux%u, %llu samples, %llu rays, %.2f ms on the GPU -> %s\n", img.width, img.height,
                (unsigned long long)st.samples, (unsigned long long)st.rays, st.ms_total, out.c_str());
    return 0;
  } catch (const Error& e) {
    std::fprintf(stderr, "error %d: %s\n", e.code, e.what());
    return e.code == RT_ERR_CUDA ? 3 : 1;
  }
}
