// compact_kernels.cuh -- particle_bcs with device-resident counts (include/cylgpu.h cylgpu_set_exchange_capacity):
// hole-filling compaction, packing of the migrants into the fixed-size message, arrivals and the window's
// removals, every count read from device memory.  Kernel-only header, included inside namespace cylgpu after
// pbcs_kernels.cuh; barrier-free, so tests/emul/ runs these very kernels on the CPU against a numpy restatement
// of "survivors + arrivals".  Replaces the list unlink / relink and the count-then-data MPI_SENDRECV pair of
// partlist_sendrecv (boundary.F90:1867-1877, partlist.F90:842,869).  Product code: no oracle here.
#pragma once

struct Soa { double* d[7]; };

#define LEAVER_GRID (148 * 2)
#define XHDR 7   // doubles in front of the payload: the count (as int64) and padding to one particle slot

// one thread: the plan of this compaction, the new count, the message headers, the statistics
__global__ void k_plan_compact(const unsigned long long* __restrict__ cnt, unsigned long long* __restrict__ cnt2,
                               int64_t* n_dev, CompactPlan* plan, int64_t* pstats, long long xcap, int window,
                               double* hdr_l, double* hdr_r) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long nholes = (long long)cnt[CNT_HOLE];
  long long nl = (long long)cnt[CNT_LEFT], nr = (long long)cnt[CNT_RIGHT];
  if (window) {
    pstats[PST_WINDOW_REMOVED] += nholes;
  } else {
    pstats[PST_SENT_L] += nl;
    pstats[PST_SENT_R] += nr;
    pstats[PST_REMOVED] += (long long)cnt[CNT_GONE];
    pstats[PST_WINDOW_REMOVED] += (long long)cnt[CNT_GONE_WINDOW];   // remove_particles riding on this classification
  }
  if (nl > xcap || nr > xcap) pstats[PST_OVERFLOW] = 1;   // the surplus is lost: reported as an error
  if (nl > xcap) nl = xcap;
  if (nr > xcap) nr = xcap;
  plan->nholes = nholes;
  plan->n_new = (long long)*n_dev - nholes;
  plan->n_left = nl;
  plan->n_right = nr;
  *n_dev = plan->n_new;
  if (hdr_l) *reinterpret_cast<long long*>(hdr_l) = nl;
  if (hdr_r) *reinterpret_cast<long long*>(hdr_r) = nr;
  for (int k = 0; k < 8; ++k) cnt2[k] = 0ULL;
}
// the tail marks of the `nholes` last slots start cleared
__global__ void __launch_bounds__(256) k_clear_tail(uint8_t* __restrict__ tailmark, const CompactPlan* __restrict__ plan) {
  const long long nholes = plan->nholes;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nholes; t += (long long)gridDim.x * blockDim.x)
    tailmark[t] = 0;
}

__global__ void __launch_bounds__(256) k_collect_dev(Soa s, const uint32_t* __restrict__ hole_list,
                                                     const uint8_t* __restrict__ hole_flag,
                                                     const CompactPlan* __restrict__ plan, double* __restrict__ send_l,
                                                     double* __restrict__ send_r, long long xcap,
                                                     uint32_t* __restrict__ lowhole, uint8_t* __restrict__ tailmark,
                                                     unsigned long long* cnt2) {
  const long long nholes = plan->nholes, n_new = plan->n_new;
  for (long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x; h < nholes; h += (long long)gridDim.x * blockDim.x) {
    const uint32_t i = hole_list[h];
    const uint8_t f = hole_flag[h];
    if (f == FL_LEFT || f == FL_RIGHT) {
      const unsigned long long k = atomicAdd(&cnt2[f], 1ULL);
      if ((long long)k < xcap) {
        double* dst = ((f == FL_LEFT) ? send_l : send_r) + 7 * k;
#pragma unroll
        for (int q = 0; q < 7; ++q) dst[q] = s.d[q][i];
      }
    }
    if ((long long)i >= n_new) tailmark[(long long)i - n_new] = 1;
    else lowhole[atomicAdd(&cnt2[4], 1ULL)] = i;
  }
}

__global__ void __launch_bounds__(256) k_tail_keepers_dev(const uint8_t* __restrict__ tailmark,
                                                          const CompactPlan* __restrict__ plan,
                                                          uint32_t* __restrict__ hightail, unsigned long long* cnt2) {
  const long long nholes = plan->nholes, n_new = plan->n_new;
  if (n_new <= 0) return;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nholes; t += (long long)gridDim.x * blockDim.x)
    if (!tailmark[t]) hightail[atomicAdd(&cnt2[3], 1ULL)] = (uint32_t)(n_new + t);
}

__global__ void __launch_bounds__(256) k_fill_holes_dev(Soa s, const uint32_t* __restrict__ lowhole,
                                                        const uint32_t* __restrict__ hightail,
                                                        const CompactPlan* __restrict__ plan,
                                                        const unsigned long long* cnt2) {
  if (plan->n_new <= 0) return;
  const long long nfill = (long long)cnt2[4];
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nfill; k += (long long)gridDim.x * blockDim.x) {
    const uint32_t dst = lowhole[k], src = hightail[k];
#pragma unroll
    for (int q = 0; q < 7; ++q) s.d[q][dst] = s.d[q][src];
  }
}

// arrivals behind the survivors: from the right neighbour first, then from the left (boundary.F90:1867-1877)
__global__ void __launch_bounds__(256) k_unpack_dev(Soa s, const int64_t* __restrict__ n_dev,
                                                    const double* __restrict__ recv_r,
                                                    const double* __restrict__ recv_l) {
  const long long from_r = recv_r ? *reinterpret_cast<const long long*>(recv_r) : 0;
  const long long from_l = recv_l ? *reinterpret_cast<const long long*>(recv_l) : 0;
  const long long base = (long long)*n_dev;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < from_r + from_l;
       k += (long long)gridDim.x * blockDim.x) {
    const double* src = (k < from_r) ? recv_r + XHDR + 7 * k : recv_l + XHDR + 7 * (k - from_r);
#pragma unroll
    for (int q = 0; q < 7; ++q) s.d[q][base + k] = src[q];
  }
}
__global__ void k_bump_count(int64_t* n_dev, int64_t* pstats, const double* recv_r, const double* recv_l) {
  const long long from_r = recv_r ? *reinterpret_cast<const long long*>(recv_r) : 0;
  const long long from_l = recv_l ? *reinterpret_cast<const long long*>(recv_l) : 0;
  *n_dev += from_r + from_l;
  pstats[PST_RECV] += from_r + from_l;
}
__global__ void k_add_count(int64_t* n_dev, long long add) { *n_dev += add; }

__global__ void __launch_bounds__(256) k_pbcs_classify_dev(BcsConst B, double* __restrict__ x, double* __restrict__ y,
                                                           double* __restrict__ z, double* __restrict__ px,
                                                           double* __restrict__ py, double* __restrict__ pz,
                                                           uint32_t* __restrict__ hole_list,
                                                           uint8_t* __restrict__ hole_flag, unsigned long long* cnt,
                                                           const int64_t* __restrict__ n_dev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_dev) return;
  const double X = x[i];
  if (X < B.remove_x) {   // behind the window that has just moved (window.F90:304-325)
    record_leaver(hole_list, hole_flag, cnt, (uint32_t)i, FL_GONE_WINDOW);
    return;
  }
  // a list that was inside before the window moved can only have left through an x face
  if (B.x_only && X >= B.x_min_local && X < B.x_max_local) return;
  MemParticle a{x + i, y + i, z + i, px + i, py + i, pz + i};
  const uint8_t f = particle_bcs_one(B, a);
  if (f != FL_KEEP) record_leaver(hole_list, hole_flag, cnt, (uint32_t)i, f);
}
__global__ void __launch_bounds__(256) k_flag_behind_dev(const double* __restrict__ x, double x_min,
                                                         uint32_t* __restrict__ hole_list,
                                                         uint8_t* __restrict__ hole_flag, unsigned long long* cnt,
                                                         const int64_t* __restrict__ n_dev) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *n_dev) return;
  if (x[i] < x_min) record_leaver(hole_list, hole_flag, cnt, (uint32_t)i, FL_GONE);
}

