// dcs_dev_probes.cuh - development probes (traffic / atomics / exchange-step timers).  NOT part of the product
// library: compiled only with -DDCS_DEV_PROBES (make dev -> libdcs_b200_dev.so), used by scripts/flat_probe.py and
// scripts/mgpu_pcg.py.  Included at the end of dcs_api.cu (needs the private handle).
#pragma once

namespace {

// development probe: K1's memory traffic as a flat, dependency-free stream (one thread per half-edge slot):
// mode bit 0: gather the other pose; bit 1: store a 3x3 block per owner slot into the compact tile-interleaved array
__global__ void k_dbg_flat(const HalfEdgeRec* __restrict__ recs, const double4* __restrict__ xyt, int64_t nslots, double* Hup, int64_t ldu, int mode) {
  const L2Policy pol = make_l2_policy();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nslots) return;
  HalfEdgeRec r;
  ld_rec(r, recs + i, pol.stream);
  double a = r.tmx, b = r.tmy, c = r.thm;
  if (mode & 1) { PoseRec p; ld_pose(p, xyt + (r.word & kIdxMask), pol.keep); a += p.x; b += p.y; c += p.th; }
  if ((mode & 2) && (r.word & kFlagOwner)) {
    double* o = Hup + block_base((i >> 1) % ldu);       // ~ the compact array's footprint
#pragma unroll
    for (int k = 0; k < kBlockVals; ++k) st_stream(o + k * 32, a + k * b + c, pol.stream);
  }
  if (!(mode & 2) && a + b + c == 1.2345e300) Hup[i] = a;
}

// development probe: scatter-add cost of an edge-centric assembly (18 fp64 reductions per edge into the
// diagonal blocks / gradient of the two endpoints)
__global__ void k_dbg_atomic(const int32_t* __restrict__ ea, const int32_t* __restrict__ eb, int32_t E, int64_t ldn, double* acc, int nper) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t a = ea[e], b = eb[e];
  for (int c = 0; c < nper; ++c) {
    atomicAdd(acc + (int64_t)c * ldn + a, 1.0 + c);
    atomicAdd(acc + (int64_t)c * ldn + b, 2.0 + c);
  }
}

}  // namespace

extern "C" {

// development probe (not part of the public header): time the flat traffic kernel; returns us per launch
DCS_API double dcs_debug_flat(dcs_handle* h, int mode, int repeats) {
  cudaSetDevice(h->dev);
  const int64_t ns = h->n_slot_tiles * kSlice;
  const int grid = cdiv(ns, 256);
  for (int i = 0; i < 3; ++i) k_dbg_flat<<<grid, 256, 0, h->stream>>>(h->recs.p, h->xyt.p, ns, h->Hup.p, h->ldu, mode);
  cudaEventRecord(h->ev0, h->stream);
  for (int i = 0; i < repeats; ++i) k_dbg_flat<<<grid, 256, 0, h->stream>>>(h->recs.p, h->xyt.p, ns, h->Hup.p, h->ldu, mode);
  cudaEventRecord(h->ev1, h->stream);
  cudaEventSynchronize(h->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  return 1e3 * ms / repeats;
}

DCS_API double dcs_debug_atomic(dcs_handle* h, int nper, int repeats) {
  cudaSetDevice(h->dev);
  double* acc = nullptr;
  cudaMalloc(&acc, (size_t)9 * h->ldn * 8);
  cudaMemset(acc, 0, (size_t)9 * h->ldn * 8);
  const int grid = cdiv(h->E, 256);
  k_dbg_atomic<<<grid, 256, 0, h->stream>>>(h->ea.p, h->eb.p, h->E, h->ldn, acc, nper);
  cudaEventRecord(h->ev0, h->stream);
  for (int i = 0; i < repeats; ++i) k_dbg_atomic<<<grid, 256, 0, h->stream>>>(h->ea.p, h->eb.p, h->E, h->ldn, acc, nper);
  cudaEventRecord(h->ev1, h->stream);
  cudaEventSynchronize(h->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  cudaFree(acc);
  return 1e3 * ms / repeats;
}

// development probe: time the exchange steps of one PCG iteration (us each, averaged over `repeats`)
DCS_API int dcs_debug_comm(dcs_handle* h, int repeats, double* out4) {
  cudaSetDevice(h->dev);
  auto timeit = [&](auto&& fn) -> double {
    for (int i = 0; i < 3; ++i) fn();
    cudaEventRecord(h->ev0, h->stream);
    for (int i = 0; i < repeats; ++i) fn();
    cudaEventRecord(h->ev1, h->stream);
    cudaEventSynchronize(h->ev1);
    float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    return 1e3 * ms / repeats;
  };
  out4[0] = timeit([&] { halo_exchange(h, h->p4.p); });
  out4[1] = timeit([&] { allreduce_sum(h, h->scal.p + S_PQ, 1); });
  out4[2] = timeit([&] { allreduce_sum(h, h->scal.p + S_TMP, 2); });
  out4[3] = (double)h->halo_send_off[h->world] * 32.0 / 1e6;   // MB sent per exchange
  return DCS_OK;
}

// development probe: per-stage device time of the PCG iteration (us, averaged), launched stage by stage with
// events in between (so launch gaps are included, unlike the CUDA-graph production path).
// out[0..5] = spmv+fold, all-reduce(p.q), vector kernel+fold, all-reduce(r.z, r.r), direction, halo exchange
DCS_API int dcs_debug_pcg_stages(dcs_handle* h, int repeats, double* out6) {
  cudaSetDevice(h->dev);
  if (!h->have_lin) return DCS_ERR_ARG;
  cudaEvent_t ev[7];
  for (auto& e : ev) cudaEventCreate(&e);
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const double* D = h->Adiag.p;
  for (int it = 0; it < repeats + 3; ++it) {
    cudaEventRecord(ev[0], h->stream);
    LAUNCH((k_spmv<double, kSpmvAll>), h->nblk, kRowsPerBlock, h->stream, h->p4.p, h->layout(), h->cols.p, h->Hoff.p, D, h->n_loc, h->q.p, h->task_part.p);
    k_fold_tasks<1, 0><<<fold_blocks(h->nblk), kFoldThreads, 0, h->stream>>>(h->task_part.p, h->nblk, h->scal.p + S_PQ, h->scal.p, 1, h->fold_ws.p, h->tickets.p + 6);
    cudaEventRecord(ev[1], h->stream);
    CKS(allreduce_sum(h, h->scal.p + S_PQ, 1));
    cudaEventRecord(ev[2], h->stream);
    k_pcg_chain<false><<<h->ntiles, kChainThreads, kChainSmemBytes, h->stream>>>((const double*)nullptr, h->is_free.p, h->p4.p, h->q.p, h->chL.p, h->chS.p,
        h->perm.p, 0, h->nrows, h->ldn, h->w.p, h->r.p, h->z.p, h->p4.p, h->task_part.p, h->scal.p);
    k_fold_tasks<2, 0><<<fold_blocks(h->ntiles), kFoldThreads, 0, h->stream>>>(h->task_part.p, h->ntiles, h->scal.p + S_TMP, h->scal.p, 0, h->fold_ws.p, h->tickets.p + 6);
    cudaEventRecord(ev[3], h->stream);
    CKS(allreduce_sum(h, h->scal.p + S_TMP, 2));
    cudaEventRecord(ev[4], h->stream);
    LAUNCH(k_pcg_direction, h->vec_grid(), kVecThreads, h->stream, h->z.p, 0, h->nrows, h->ldn, h->p4.p, h->scal.p);
    cudaEventRecord(ev[5], h->stream);
    CKS(halo_exchange(h, h->p4.p));
    cudaEventRecord(ev[6], h->stream);
    cudaEventSynchronize(ev[6]);
    if (it >= 3)
      for (int s = 0; s < 6; ++s) { float ms = 0; cudaEventElapsedTime(&ms, ev[s], ev[s + 1]); acc[s] += 1e3 * ms; }
  }
  for (int s = 0; s < 6; ++s) out6[s] = acc[s] / repeats;
  for (auto& e : ev) cudaEventDestroy(e);
  return DCS_OK;
}

// development probe: per-phase cycle counts of the last k_pcg_cluster solve on this device (see dcs_pcg_cluster.cuh)
DCS_API int dcs_debug_cluster_cycles(double* out8) {
  unsigned long long c[8];
  if (cudaMemcpyFromSymbol(c, dcs::g_cl_cycles, sizeof(c)) != cudaSuccess) return DCS_ERR_CUDA;
  for (int i = 0; i < 8; ++i) out8[i] = (double)c[i];
  return DCS_OK;
}

}  // extern "C"
