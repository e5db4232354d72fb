// sixteen_armadillos.cpp -- the reference's examples/sixteen_armadillos.rs written against the C++ host mirror
// (bvhtracer_b200/host/bvhtracer.hpp).  Same scene, camera, closed-form animation, accumulator and pixel shader; the only
// difference to the Rust example is `CudaPathTracer` where it says `PathTracer` (sixteen_armadillos.rs:184).
//
// Build and run (tests/test_cpp_example.py does exactly this):
//   g++ -O2 -std=c++17 -ffp-contract=off examples/sixteen_armadillos.cpp -Lbvhtracer_b200/lib -lbvht_cuda -Wl,-rpath,<lib dir> -o sixteen
//   ./sixteen assets/armadillo.tri.f32 30
#include <chrono>
#include <cstdio>
#include <fstream>

#include "../bvhtracer_b200/host/bvhtracer.hpp"

using namespace bvhtracer;

constexpr size_t SCREEN_WIDTH = 640, SCREEN_HEIGHT = 640;          // sixteen_armadillos.rs:6-7

struct Physics {                                                    // sixteen_armadillos.rs:21-44
    float angle; Vector3 position_init; float height, speed, angular_velocity, acceleration, direction = -1.0f;
};

static Mesh load_packed_tri(const char* path) {                     // assets/*.f32: N x 9 f32 (oracle/tools/pack_assets.py)
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    size_t bytes = (size_t)f.tellg();
    f.seekg(0);
    std::vector<float> v(bytes / 4);
    f.read((char*)v.data(), (std::streamsize)bytes);
    MeshBuilder b;
    for (size_t k = 0; k + 8 < v.size(); k += 9) {
        Triangle t;
        for (int i = 0; i < 3; ++i) t.vertices[i] = Vector3(v[k + 3 * i], v[k + 3 * i + 1], v[k + 3 * i + 2]);
        b.with_primitive(t);
    }
    return b.build();
}

int main(int argc, char** argv) {
    const char* asset = argc > 1 ? argv[1] : "assets/armadillo.tri.f32";
    int frames = argc > 2 ? std::atoi(argv[2]) : 10;
    try {
        Camera camera(SymmetricFovSpec{ 90.0f, 1.0f, 2.0f, 10000.0f },
                      CameraAttitudeSpec{ Vector3(0, 1, -5.5f), Vector3::unit_z(), Vector3::unit_x(), Vector3::unit_y(), Vector3::unit_z() });
        ModelInstance model = ModelBuilder().with_mesh(load_packed_tri(asset)).build();
        std::vector<SceneObject> objects;
        std::vector<Physics> physics;
        const float height_init[16] = { 5, 4, 3, 2, 1, 5, 4, 3, 5, 4, 3, 2, 1, 5, 4, 3 };
        int i = 0;
        for (int x = 0; x < 4; ++x) for (int y = 0; y < 4; ++y) {
            bool even = ((x + y) & 1) == 0;
            float angular_velocity = even ? (float)((i * 13) & (7 + 2)) * 0.10f : 0.0f;      // `&` binds weaker than `+` in Rust
            Vector3 horizontal(((float)x - 1.5f) * 2.5f, 0.0f, ((float)y - 1.5f) * 2.5f);
            Vector3 vertical = even ? Vector3::zero() : Vector3(0.0f, height_init[i / 2], 0.0f);
            Vector3 translation = horizontal + vertical;
            Transform3 transform = Transform3::new_(Vector3::from_fill(0.75f), translation, Rotation3::from_angle_x(0.0f) * Rotation3::from_angle_z(0.0f));
            physics.push_back(Physics{ 0.0f, translation, vertical.y, 0.0f, angular_velocity, even ? 0.0f : 9.8f });
            objects.push_back(SceneObjectBuilder(model).with_transform(transform).build());
            ++i;
        }
        Scene scene = SceneBuilder(camera).with_objects(std::move(objects)).build();
        Renderer renderer(std::make_unique<CudaPathTracer>(BVHT_FLAG_STRICT | BVHT_FLAG_LEAF_ACCEL));
        RendererState state(ShadingPipeline::depth(80.0f, 3.0f), SCREEN_WIDTH, SCREEN_HEIGHT, false);

        auto t0 = std::chrono::steady_clock::now();
        size_t rays = renderer.render(state, scene);
        const float elapsed = (float)(1.0 / 60.0);
        for (int f = 0; f < frames; ++f) {
            for (int k = 0; k < 16; ++k) {                                                    // AppState::update, :132-163
                Vector3 translation = physics[k].position_init + Vector3(0.0f, physics[k].height, 0.0f);
                Rotation3 rotation = Rotation3::from_angle_x(physics[k].angle) * Rotation3::from_angle_z(physics[k].angle);
                scene.get_mut_unchecked(k).set_transform(Transform3::new_(Vector3::from_fill(0.75f), translation, rotation));
            }
            for (int k = 0; k < 16; ++k) {
                Physics& p = physics[k];
                p.angle = p.angle + p.angular_velocity * elapsed;
                p.speed += p.acceleration * elapsed;
                p.height += p.direction * p.speed * elapsed;
                if (p.height < -3.0f) { p.height = -3.0f + 0.01f; p.direction = -p.direction; p.speed = 0.2f; }
                else if (p.height > p.position_init.y) { p.height = p.position_init.y - 0.01f; p.direction = -p.direction; }
            }
            scene.rebuild();
            rays += renderer.render(state, scene);
        }
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        uint32_t checksum = 0;
        for (size_t p = 0; p < SCREEN_WIDTH * SCREEN_HEIGHT; ++p) checksum ^= state.frame_buffer()[p] * (uint32_t)(p | 1);
        std::printf("sixteen_armadillos: %d frames, %zu rays, %.2f ms, frame checksum %08x\n", frames + 1, rays, ms, checksum);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
