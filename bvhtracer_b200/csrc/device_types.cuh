// device_types.cuh -- HBM-resident layouts shared by the upload, refit and trace kernels.
//
// Everything here is the UPLOADED form of the reference's in-memory objects (DESIGN.md "Data layout"):
//   BvhNode  (bvh.rs:88-134, 32 B)  -> 2 x float4 : {min.xyz, left_first}, {max.xyz, prim_count}
//   Triangle (triangle.rs:9-16, 36 B) -> 3 float4 per triangle: v0 | e1 = v1 - v0 | e2 = v2 - v0  (48 B, 16 B aligned)
//            (one IEEE subtraction each, exactly what Triangle::intersect computes first, triangle.rs:43-44)
//   TlasNode (tlas.rs:41-46, 32 B)  -> 2 x float4 : {min.xyz, left_right}, {max.xyz, blas}
//   SceneObject inverse transform (scene_object.rs:78-89) -> 4 x float4 columns
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bvht {

// Form of the baked (inflated) sub-BVH child boxes: 0 = {lo, hi}, 1 / 2 = {centre, half extent} (trace_kernels.cuh
// slab_test_ch; 2 keeps |f| in registers and uses three FMAs per axis).  Shared by the bake and the trace kernels.
#ifndef BVHT_SUB_CH
#define BVHT_SUB_CH 2
#endif

// Chain-skipping TLAS walk for blocks with a candidate mask (trace_kernels.cuh build_tlas_skip): 1 = on
#ifndef BVHT_TLAS_PRUNE
#define BVHT_TLAS_PRUNE 1
#endif

constexpr int kTlasStack = 64;   // validated at bvht_tlas_set (depth of the uploaded tree)
constexpr int kBlasStack = 64;   // validated at bvht_blas_create
constexpr int kSubStack  = 48;   // leaf sub-BVH (built by us, depth bounded at build)

// One per uploaded model.  Pointers are device addresses.
struct BlasDesc {
    const float4* nodes;      // 2 float4 per reference node
    const float4* tri;        // 3 float4 per triangle, REFERENCE primitive order: {v0, prim}, {e1, 0}, {e2, 0}   (48 B, one
                              //   base address + immediate offsets per test; brute-force leaves and refit source)
    // leaf accelerator (BVHT_FLAG_LEAF_ACCEL), see leaf_accel.hpp
    const float4* sub_nodes;  // 4 float4 per sub node (two child boxes + two child refs), inflated for the current bake
    const float4* stri;       // 3 float4 per triangle in sub-BVH order; v0.w carries the reference primitive index
    const uint32_t* leaf_sub_root;  // per reference node: sub-BVH root ref for leaves (0xFFFFFFFF = brute force)
    uint32_t n_tris;
    uint32_t nodes_used;
    float    accel_d_max;     // model-space limits |d| <= d_max, |o| <= o_max under which the inflated
    float    accel_o_max;     //   sub boxes are conservative (leaf_accel.hpp); other rays use brute-force leaves
    float    bake_scale;      // the current bake's inflation of a triangle's box: scale * |e1||e2| + abs (accel_deltas)
    float    bake_abs;
};

struct SceneDev {
    const float4*   tlas_tight;  // accel only, 2 float4 per TLAS node: conservative world box of the REAL geometry below the
                                 //   node {lo.xyz, max |d_w|^2} {hi.xyz, max |o_w - tight_center|^2} (limits under which it may be used)
    const uint32_t* tlas_mask;   // accel only, n_inst <= 32: bit i set when instance i is below the node
    const float4*   tlas;        // 2 float4 per TLAS node
    const float4*   inst_cols;   // 4 float4 per instance: columns of the inverse transform
    const uint32_t* inst_blas;   // blas id per instance
    const BlasDesc* blas;
    uint32_t        n_inst;
    uint32_t        flags;
    float           tight_center[3];   // the tight boxes' origin limit is |o_w - tight_center|^2 <= hi.w
};

struct CameraDev {
    float tl[3], tr[3], bl[3];
    float vinv[16];              // column-major
};

struct PrimaryParams {
    SceneDev  scene;
    CameraDev cam;
    uint32_t  width, height, tile;
    uint32_t  x0, y0, x1, y1;            // pixel region
    uint32_t  tx0, ty0, ntx, nty;        // tile range covering the region (nty counts only this shard's tile rows)
    uint32_t  row_stride;                // tile-row interleave across GPUs: this launch owns rows ty0 + k * row_stride
    uint32_t  items_per_tile;            // ceil(tile*tile / 32)
    uint32_t  n_items;                   // ntx * nty * items_per_tile
    uint4*    out;                       // width*height hit records (16 B each); may be null when out_rgba is set
    uint32_t* out_rgba;                  // width*height Rgba<u8> pixels (fused accumulator + pixel shader); may be null
    uint32_t  shade_kind;                // bvht_shade_kind
    float     shade_scale, shade_offset; // DepthMappingShader::new(scale, offset)
    uint32_t  hit_rgba, miss_rgba;       // IntersectionShader::new(hit, miss), packed r | g<<8 | b<<16 | a<<24
    const float4* shade_normals;         // kind 4: 3 float4 per primitive of scene object 0's model (un-reordered normals)
    uint32_t  shade_n_prims;
    const float2* shade_tex;             // kind 5: 3 float2 per primitive of scene object 0's model (un-reordered tex coords)
    const uint8_t* shade_texels;         // kind 5: object 0's texture, Rgb<u8>, texel (x, y) at (y * tex_w + x) * 3
    uint32_t  tex_w, tex_h;
    float     shade_m[12];               // kind 4: columns 0..2 (xyz) + column 3 (xyz) of object 0's forward transform
    unsigned int* work_counter;          // persistent-thread work cursor
    // accel + candidate masks: K0 (classify_fill_kernel) writes the records of the blocks no instance can be seen from and
    // lists the others here; K1 then pulls list positions instead of block numbers.  null = K1 walks every block itself.
    uint32_t*     work_list;
    // Bands.  The launch's pixel blocks, in their row-major numbering, are cut into n_bands runs of band_items blocks (whole tile
    // rows; the last run may be shorter).  K1 pulls them band by band in band_order -- the host puts the cheapest bands first,
    // so that the frame ends with its most expensive rows and the device->host copies of the earlier bands hide behind them --
    // and raises band_flag[b] to band_seq when the last block of band b is finished (system-scope fence first: the flags of a
    // host-bound frame live in page-locked host memory, and the host issues the copy of that band's rows when it sees the flag;
    // bvht_api.cu pump_flights), with no kernel boundary between bands.
    uint32_t      n_bands;               // 1..32
    uint32_t      band_items;            // blocks per band = rows per band * ntx * items_per_tile
    uint8_t       band_order[32];        // pull position -> band
    unsigned int* band_count;            // [32] blocks K0 listed per band (zeroed before K0); with a work list, band b's entries
                                         //   are work_list[b * band_items .. + band_count[b])
    unsigned int* band_done;             // [32] blocks finished per band (zeroed per frame); null = no completion flags
    unsigned int* band_flag;             // [32]
    uint32_t      band_seq;
    unsigned long long* stats;           // debug counters (stats build only)
    // accel only: conservative screen-space rectangle (pixels, inclusive) of each instance's tight box for THIS camera;
    // n_rect == 0 disables the tile-level candidate masks
    // the camera origin in every instance's model space (primary rays share it): computed on the host with the kernel's own
    // operation order (BVHT_MV4, no contraction), so the kernel need not transform the origin at every instance entry
    uint32_t  n_origin;                  // 0 = not available (more than 32 instances)
    float4    inst_origin[32];
    uint32_t  n_rect;
    // per-triangle coverage of the 8x4 blocks (cover_kernels.cu; tile == 8 only): bit i of cover[(ty * cover_ntx + tx) * 2 + half]
    // = some triangle of instance i may be seen from that block; *cover_full = instances to be treated as visible everywhere
    const uint32_t* cover;               // null = not available
    const uint32_t* cover_full;
    uint32_t  cover_ntx;
    uint32_t  skip_rounds;               // pointer-jumping rounds: ceil(log2(depth of the TLAS))
    uint32_t  n_tlas_nodes;              // > 0 (and <= 64) with n_rect: chain-skipping TLAS walk (trace_kernels.cuh build_tlas_skip)
    int4      inst_rect[32];
};

// cover_kernels.cu: rasterise every instance's triangles (conservative boxes) onto the 8x4-pixel blocks of the frame
struct CoverParams {
    uint32_t*       cover;          // one word per block, zeroed before the launch
    uint32_t*       full;           // one word, zeroed before the launch
    uint32_t        full_init;      // instances the host already knows to be visible everywhere (OR-ed into *full by the kernel)
    uint32_t*       big_count;      // one word, zeroed
    int4*           big_list;       // big_cap entries: {instance, bx0 << 16 | bx1, by0 << 16 | by1, -}
    uint32_t        big_cap;
    const BlasDesc* blas;
    const uint32_t* inst_blas;
    uint32_t        n_inst;         // <= 32
    uint32_t        tri_offset[33]; // prefix sums of the instances' triangle counts
    uint32_t        ntx;            // tiles (= blocks) per row of the full frame
    uint32_t        shard_index, shard_count;   // multi-GPU: only the tile rows r with r % shard_count == shard_index are marked
    uint32_t        width, height;
    float           tlx, tly, inv_ex, inv_ey, near_, z_eps;     // near-plane rectangle of the camera (eye space)
    float           mv[32][12];     // per instance: view * object transform, rows 0..2 (row-major 3 x 4)
    const float4*   inst_tri[32];   // per instance: its model's triangle records and the current bake's inflation, copied from the
    float           inst_scale[32]; //   BLAS descriptors by the host (two dependent global loads less per thread of a kernel that is
    float           inst_abs[32];   //   bound by exactly such latencies)
};

// K6 (scene_kernels.cu): SceneObject::set_transform for every object + Tlas::rebuild
struct SceneRebuildParams {
    const float*    transforms;   // n_inst x 16, column-major FORWARD transforms (Transform3, transform.rs:24-43)
    const uint32_t* blas_ids;     // n_inst
    const BlasDesc* blas;
    float4*         tlas;         // out: 2 * n_inst nodes, 2 float4 each
    float4*         inst_cols;    // out: inverse transforms, 4 float4 per instance
    uint32_t*       inst_blas;    // out: copy of blas_ids
    float*          inst_bounds;  // out (may be null): world AABB per instance, 6 floats (SceneObject::bounds)
    unsigned int*   status;       // [0]: bit 0 = singular transform, bit 1 = clustering found no candidate; [1] = nodes_used
    void*           pack;         // out (may be null): tlas | inst_cols | inst_bounds | status in one contiguous block (one D2H)
    uint32_t        n_inst;
};

struct RaysParams {
    SceneDev     scene;
    const float* rays;                   // n x 7 floats (o, d, t)
    uint64_t     n;
    uint4*       out;
    unsigned int* work_counter;
};

} // namespace bvht
