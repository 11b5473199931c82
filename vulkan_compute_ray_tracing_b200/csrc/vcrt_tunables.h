// vcrt_tunables.h -- compile-time knobs of the kernels (override with make EXTRA=-D...; each default is the measured best).
#pragma once

#define VCRT_BLOCK 128
#ifndef VCRT_PBLOCK
#define VCRT_PBLOCK 128   /* persistent kernel: threads per block */
#endif
#ifndef VCRT_PMINB
#define VCRT_PMINB 10    /* persistent kernels: min resident blocks per SM = register cap 48 (r01 A/B on C3: 1 -> 4328, 9 -> 4487, 10 -> 4586, 12 -> 4516 Mrays/s) */
#endif

#ifndef VCRT_MEGA_MINB
#define VCRT_MEGA_MINB 1  /* megakernel (A/B variant): it carries the shading state too, no register cap */
#endif
#ifndef VCRT_VISITS
#define VCRT_VISITS 2  /* trace kernel: inner-node visits per lane between two warp votes on the phase switch */
#endif
#ifndef VCRT_VISITS4
#define VCRT_VISITS4 1 /* the same for 4-wide nodes (r01 A/B on C3: 1 -> 6501, 2 -> 6224 Mrays/s in the trace kernel) */
#endif
#ifndef VCRT_SMEMRAY
#define VCRT_SMEMRAY 1   /* trace kernel: ray origin/direction parked in shared memory between triangle tests: no spills at 48 registers (r01 A/B on C3, 4-wide nodes: 6482 -> 6657 Mrays/s; more resident blocks did not pay: 12 blocks 6311) */
#endif
#ifndef VCRT_SSTACK
#define VCRT_SSTACK 0    /* trace kernel: traversal-stack slots per lane kept in shared memory, [slot][thread] (deeper slots: local memory); r01 A/B on C3: 8/16/24 slots all 3-4 % slower than local memory */
#endif
#ifndef VCRT_SHADE_MINB
#define VCRT_SHADE_MINB 1  /* shade kernel: min resident blocks (of 256) per SM; 1 = no register cap (60 registers).  r01 A/B on C3 at 32 spp, whole pipeline: 1 -> 5649, 5 (48 registers) -> 5522, 6 -> 5445, 8 -> 5386 Mrays/s */
#endif
#ifndef VCRT_SHADE_GRID
#define VCRT_SHADE_GRID 8u /* shade kernel: blocks per SM in the grid-stride launch (8 vs 20: no difference) */
#endif
#ifndef VCRT_PF_L2
#define VCRT_PF_L2 0     /* trace kernel, 4-wide nodes: prefetch.global.L2 of the parked leaf's triangle and of the new top-of-stack child (A/B for HBM-sized scenes) */
#endif
#ifndef VCRT_PREFETCH
#define VCRT_PREFETCH 0  /* trace kernel: 1 = prefetch the triangle of a postponed leaf into L1 (measured: 32 % SLOWER on C3, r01) */
#endif

#ifndef VCRT_WF_AUTO2_PATHS
#define VCRT_WF_AUTO2_PATHS (1ull << 25)  /* wf_streams=auto: calls of this many paths or more run as two pipelines (vcrt_api.cu).  C3, 1 vs 2 pipelines: 4 spp 4945 / 4876, 8 spp 5725 / 5732, 16 spp (33.2 M paths) 6231 / 6286, 64 spp 6436-6651 / 6763-6765 Mrays/s (profiles/r02_v55_ab_pipelines.log) */
#endif

#ifndef VCRT_TAIL_SPLIT
#define VCRT_TAIL_SPLIT 1  /* trace kernel: once the queue is dry, idle lanes of a warp take subtrees off the stacks of its busy lanes (vcrt_wavefront.cuh) */
#endif
#ifndef VCRT_TAIL_ENTER
#define VCRT_TAIL_ENTER 16  /* ... the tail loop takes over once fewer than this many lanes of the warp are still walking (33 = as soon as the queue is dry).  33 / 28 / 24 / 16: C3 at 1 spp 2581 / 2578 / 2576 / 2629 Mrays/s, 8 and 64 spp unchanged (profiles/r02_v46_ab_tail_enter.log) */
#endif
#ifndef VCRT_TAIL_EXCHANGE
#define VCRT_TAIL_EXCHANGE 0  /* ... owners and helpers exchange their closest hits at every round of reports, not only at the end: the loop over all helpers costs more than the culling returns -- C3 at 1 / 2 / 8 spp 2584 -> 2444, 3764 -> 3600, 5611 -> 5525 Mrays/s (profiles/r02_v37_ab_tail_exchange.log) */
#endif
#ifndef VCRT_TAIL_PREFETCH
#define VCRT_TAIL_PREFETCH 0  /* ... the records of the children a ray enters are prefetched into L1 (tail loop only): slower even there -- C3 at 1 / 8 spp 2567 -> 2228, 5631 -> 5326 Mrays/s, the 10 M-triangle scene 4374 -> 4178 (profiles/r02_v28_ab_tail_prefetch.log) */
#endif
#ifndef VCRT_TAIL_ROUNDS
#define VCRT_TAIL_ROUNDS 2  /* ... visit / leaf rounds between two rounds of reports and donations (1 / 2 / 4 rounds x 1 / 2 / 4 passes: all within 3 %, profiles/r02_v26_ab_tail_split_tuning.log) */
#endif
#ifndef VCRT_TAIL_PASSES
#define VCRT_TAIL_PASSES 1  /* ... donation passes per round of donations (a donor gives one subtree per pass) */
#endif
/* tail loop, also tried: the loads of a lane's node and of its postponed triangle issued together (one round trip per round instead of two): ~100 bytes of
 * spills in the tail loop, C3 at 1 / 2 / 8 spp 2597 -> 2312, 3796 -> 3477, 5652 -> 5436 Mrays/s (profiles/r02_v33_ab_tail_combined_loads.log) */
/* r02 experiments that did not make it (logs under profiles/, code under tools/experiments/):
 *   cache policy of the triangle-record loads (L2 evict-first / no hint / L1 no-allocate instead of L2 evict-last): +-1 % on C3 and on the
 *     10 M-triangle scene, L1 no-allocate -3 % (r02_v12_ab_tri_policy.log);
 *   ray binning -- the rays of the next bounce queued in the order of the grid cell (8^3..8^6 cells) that holds their origin, filed by a
 *     warp-aggregated atomic per cell and placed directly by the shade kernel, no sort and no extra pass over the rays: the bounce
 *     launches get 5-6 % faster on C3 (bounce 1, whose rays leave the shared primary hit points, gains nothing; bounce 7: -13 %) and 7 % on
 *     the 10 M-triangle scene, but filing, scan and the shade kernel's scattered 48-byte stores cost more than that: 11.7 vs 11.1 ms at
 *     16 spp, 43.3 vs 40.3 ms at 64 spp (r02_v21_ab_ray_binning.log).  A full radix sort of the queue (cub, 30-bit Morton key + a pass
 *     that moves the rays) makes the bounce launches 15-19 % faster and the step 12 % slower (r02_v14_ab_ray_sort_potential.log). */
