// rt_kernels.h — launch interface between the C ABI (rt_api.cu) and the kernels (rt_kernels.cu).
#ifndef RT_KERNELS_H
#define RT_KERNELS_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rt_b200.h"
#include "rt_types.h"

namespace rt {

// ray queue, SoA of quads (48 B per path): A=(o.xyz,d.x) B=(d.yz,T.x,T.y) C=(T.z,pixel,sample|bounce<<24,_)
struct rt_paths {
  float4* A;
  float4* B;
  float4* C;
};
// hit records written by k_trace for k_sort / k_shade (20 B per ray): H=(t,u,v,triangle id) obj=object or -1
struct rt_hits {
  float4* H;
  int32_t* obj;
};
// secondary-ray sort: key per next-queue position, the resulting visiting order, bin counts and cursors
struct rt_sortbuf {
  uint32_t* keys;
  uint32_t* order;
  uint32_t* hist;
  uint32_t* cursor;
  uint32_t* slice_total;  // one per 1024 bins
};
// parity hooks only: resolved surface per ray: S0=(hitpoint.xyz,n.x) S1=(n.yz,u,v) S2=class|frontface<<3|id<<4
struct rt_debug {
  float4* S0;
  float4* S1;
  uint32_t* S2;
};

void launch_init(rt_ctrl* ctrl, unsigned long long begin, unsigned long long end, cudaStream_t st);  // work indices [begin, end)
void launch_advance(rt_ctrl* ctrl, uint32_t capacity, cudaStream_t st);
void launch_raygen(const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, cudaStream_t st);
void launch_trace(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_hits hits, rt_sortbuf sort,
                  bool count, uint32_t persistent_blocks, cudaStream_t st);
void launch_raysort(const rt_frame& fr, rt_ctrl* ctrl, rt_sortbuf sort, cudaStream_t st);
void launch_sort(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_hits hits, uint32_t* queues, cudaStream_t st);
void launch_surface(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_hits hits, rt_debug dbg,
                    cudaStream_t st);
int trace_blocks_per_sm();
// megakernel engine: one persistent kernel renders work indices [ctrl->cursor, ctrl->total) into the accumulator
int path_blocks_per_sm();
void launch_path(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, long long* accum, unsigned long long total,
                 uint32_t persistent_blocks, cudaStream_t st);
void launch_shade(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_paths nxt, rt_hits hits,
                  const uint32_t* queues, long long* accum, rt_sortbuf sort, bool count, cudaStream_t st);
void launch_set_window(rt_ctrl* ctrl, uint32_t n_cont, uint32_t n_new, cudaStream_t st);
void launch_phong_primary(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_paths nxt, rt_hits hits,
                          cudaStream_t st);
void launch_phong_shadow(const rt_dev_scene& sc, const rt_frame& fr, rt_ctrl* ctrl, rt_paths cur, rt_hits hits, long long* accum,
                         cudaStream_t st);
void launch_resolve(const long long* accum, uint32_t npix, uint32_t spp, float gamma, float* out_linear,
                    uint8_t* out_rgb8, cudaStream_t st);

}  // namespace rt
#endif
