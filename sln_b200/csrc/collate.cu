// GPU-side batch assembly: the step immediately before the VAE path.
// Reference: data/suncg_dataset.py:295-337 (suncg_collate_fn: concatenate the scenes of a batch, offset the subject/object ids of
// every triple by the number of objects that precede its scene, tag objects / triples with their scene's position in the batch)
// followed by utils.py:114-124 (tensor_aug: eight separate .cuda() copies).
//
// Here the host packs the scenes of a batch into ONE pinned "wire" buffer that already has the flat layout of the outputs
// (objs | angles | attributes | local triples | boxes | room ids, preceded by the per-scene prefix sums), one H2D copy moves
// it, and k_collate_finish does the per-triple / per-object work on the device.  objs, angles, attributes, boxes and ids are
// used in place (views of the device wire buffer); only triples, obj_to_img and triple_to_img are produced.
#include "../../include/sln_b200.h"
#include "common.cuh"

namespace sln {
namespace {

struct WireLayout { int64_t scene_index, obj_off, tri_off, ids, objs, angles, attrs, triples, boxes, total; };

// byte offsets inside the wire buffer; every section is 16-byte aligned
WireLayout wire_layout(int64_t B, int64_t O, int64_t T, int32_t box_dim) {
  WireLayout w;
  int64_t at = 0;
  auto take = [&](int64_t bytes) { int64_t r = at; at += (bytes + 15) / 16 * 16; return r; };
  w.scene_index = take(8 * B);
  w.obj_off = take(8 * (B + 1));
  w.tri_off = take(8 * (B + 1));
  w.ids = take(8 * B);
  w.objs = take(8 * O);
  w.angles = take(8 * O);
  w.attrs = take(8 * O);
  w.triples = take(8 * 3 * T);
  w.boxes = take(4 * (int64_t)box_dim * O);
  w.total = at;
  return w;
}

// index of the scene whose [off[i], off[i+1]) range holds x   (off has B+1 ascending entries, off[0] = 0)
__device__ __forceinline__ int find_scene(const int64_t* __restrict__ off, int B, int64_t x) {
  int lo = 0, hi = B;   // invariant: off[lo] <= x < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// thread i < T: triple i;  T <= i < T + O: object i - T.   err: [0] = number of triples whose local ids fall outside their scene
__global__ void __launch_bounds__(256) k_collate_finish(const int64_t* __restrict__ scene_index, const int64_t* __restrict__ obj_off,
                                                        const int64_t* __restrict__ tri_off, const int64_t* __restrict__ tri_local,
                                                        int B, int64_t O, int64_t T, int64_t* __restrict__ triples,
                                                        int64_t* __restrict__ obj_to_img, int64_t* __restrict__ triple_to_img,
                                                        int* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < T) {
    const int sc = find_scene(tri_off, B, i);
    const int64_t base = __ldg(obj_off + sc), n = __ldg(obj_off + sc + 1) - base;
    const int64_t s = __ldg(tri_local + 3 * i), p = __ldg(tri_local + 3 * i + 1), o = __ldg(tri_local + 3 * i + 2);
    if (err && (s < 0 || s >= n || o < 0 || o >= n)) atomicAdd(err, 1);
    triples[3 * i] = s + base;           // suncg_dataset.py:318-320
    triples[3 * i + 1] = p;
    triples[3 * i + 2] = o + base;
    if (triple_to_img) triple_to_img[i] = __ldg(scene_index + sc);   // :324
  } else if (i < T + O) {
    const int64_t j = i - T;
    if (obj_to_img) obj_to_img[j] = __ldg(scene_index + find_scene(obj_off, B, j));   // :323
  }
}

}  // namespace
}  // namespace sln

using namespace sln;

extern "C" {

int sln_collate_layout(int64_t B, int64_t O, int64_t T, int32_t box_dim, int64_t* offsets10) {
  SLN_CHECK_ARG(B >= 0 && O >= 0 && T >= 0 && box_dim > 0 && offsets10, "collate_layout: bad extents / null output");
  WireLayout w = wire_layout(B, O, T, box_dim);
  const int64_t v[10] = {w.scene_index, w.obj_off, w.tri_off, w.ids, w.objs, w.angles, w.attrs, w.triples, w.boxes, w.total};
  for (int i = 0; i < 10; ++i) offsets10[i] = v[i];
  return SLN_OK;
}

int sln_collate_finish(const void* wire, size_t wire_bytes, int64_t B, int64_t O, int64_t T, int32_t box_dim, int64_t* triples,
                       int64_t* obj_to_img, int64_t* triple_to_img, int32_t* err_count, void* stream) {
  SLN_CHECK_ARG(B >= 0 && O >= 0 && T >= 0 && box_dim > 0, "collate_finish: bad extents");
  SLN_CHECK_ARG(B < (1 << 30), "collate_finish: too many scenes");
  WireLayout w = wire_layout(B, O, T, box_dim);
  SLN_CHECK_ARG(wire && wire_bytes >= (size_t)w.total, "collate_finish: wire buffer too small (%zu < %lld)", wire_bytes, (long long)w.total);
  SLN_CHECK_ARG((uintptr_t)wire % 16 == 0, "collate_finish: wire buffer must be 16-byte aligned");
  SLN_CHECK_ARG(T == 0 || triples, "collate_finish: null triples output");
  if (O + T == 0 || B == 0) return SLN_OK;
  const char* base = (const char*)wire;
  const int64_t n = O + T;
  k_collate_finish<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const int64_t*)(base + w.scene_index), (const int64_t*)(base + w.obj_off), (const int64_t*)(base + w.tri_off),
      (const int64_t*)(base + w.triples), (int)B, O, T, triples, obj_to_img, triple_to_img, err_count);
  return check_launch("collate_finish");
}

}  // extern "C"
