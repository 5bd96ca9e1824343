// emit_runs.cuh -- reduce-side emit for fixed-framing runs read in place (run-table mode, sorter_kernels.cuh RunTable):
// the merge of G sorted runs takes CONSECUTIVE records from each run, so the records of one output tile are at most G
// contiguous byte ranges of the input segments -- and an input record already carries the framing bytes the output
// needs (vint(klen) vint(vlen) key value), so the output body is a permutation of whole input records.
//
//   producers (ET_RUN_PW warps, alternating tiles): sorted order -> record index -> run; per-run [first, last] record of
//     the tile; ONE cp.async.bulk.shared::cluster.global per run moves the run's byte range (16-byte aligned superset)
//     into the stage -- <= G bulk copies per 256-record tile instead of 256 x six 128-bit gather loads plus four SHFL
//     per piece (emit_pipe_u.cuh); the stage's mbarrier flips when the bytes have landed;
//   consumers (8 warps = the 256-thread chunk interleave of the checksum): every aligned 16-byte chunk of the output
//     image is assembled from the one or two staged records it covers (two LDS.128 + funnel shift each), folded into the
//     tile's CRC32 and streamed to HBM; no image buffer, no CTA-wide barrier per tile.
// Same tiles, byte-exact output and per-tile checksum algebra as the other emit kernels.
#pragma once
#include "emit_tma.cuh"

#ifndef TEZGPU_EMIT_RUNS_MIN_CTAS
#define TEZGPU_EMIT_RUNS_MIN_CTAS 2
#endif

namespace tezgpu {

constexpr int ER_PW = 2;                          // producer warps (tile parity)
constexpr int ER_CW = FE_THREADS / 32;            // consumer warps
constexpr int ER_THREADS = 32 * (ER_PW + ER_CW);
constexpr int ER_STAGES = 3;
constexpr int ER_BATCH = ER_CW;
constexpr int ER_MAX_RUNS = 32;                   // runs per partition one warp can plan (lane g <-> run g)

struct RunsTileGeom {
  const uint32_t *src;   // shared: src[j] = shared address of record j's first byte (its framing)
  uint32_t rec0;         // image offset of the first record
  uint32_t nr, body;     // records, nr * rec_size
  bool first, last;
};

// one byte of the tile's output image at offset rel from the first record
template <class S>
__host__ __device__ __forceinline__ uint32_t runs_image_byte(const S &sm, uint32_t rec_size, uint32_t magic, const RunsTileGeom &g, int32_t rel) {
  if (rel < 0) {
    if (!g.first || rel < -4) return 0u;
    return rel == -4 ? (uint32_t)'T' : (rel == -3 ? (uint32_t)'I' : (rel == -2 ? (uint32_t)'F' : 0u));
  }
  if ((uint32_t)rel >= g.body) return (g.last && (uint32_t)rel < g.body + 2u) ? 0xFFu : 0u;
  const uint32_t j = mulhi32((uint32_t)rel, magic), w = (uint32_t)rel - j * rec_size;
  return sm.ld8(g.src[j] + w);
}

// the aligned 16-byte chunk of the output image at image offset X
template <class S>
__host__ __device__ __forceinline__ uint4 runs_assemble(const S &sm, uint32_t rec_size, uint32_t magic, const RunsTileGeom &g, uint32_t X) {
  const int32_t rel = (int32_t)X - (int32_t)g.rec0;
  if (rel >= 0 && (uint32_t)rel + 16u <= g.body) {
    const uint32_t j = mulhi32((uint32_t)rel, magic), w = (uint32_t)rel - j * rec_size;
    const uint4 a = lds16_any(sm, g.src[j] + w);
    const uint32_t n1 = rec_size - w;             // bytes of record j from w on
    if (n1 >= 16u) return a;
    uint4 acc = low_bytes(a, n1);
    uint32_t t = n1, jj = j + 1;
    // the rest comes from the following record(s) (one, unless records are shorter than 16 bytes)
    while (t < 16u) {
      const uint32_t n = rec_size < 16u - t ? rec_size : 16u - t;
      acc = or4(acc, shl_bytes(low_bytes(lds16_any(sm, g.src[jj]), n), t));
      t += n;
      jj++;
    }
    return acc;
  }
  uint32_t w4[4] = {0, 0, 0, 0};
  for (uint32_t b = 0; b < 16u; b++) w4[b >> 2] |= runs_image_byte(sm, rec_size, magic, g, rel + (int32_t)b) << (8u * (b & 3u));
  return make_uint4(w4[0], w4[1], w4[2], w4[3]);
}

struct EmitRunsLayout {
  static constexpr size_t BARS = 256;
  static constexpr size_t TABS = 256 * 4 + 4 * 256 * 4;
  // a stage: the runs' byte ranges (each a 16-byte aligned superset, packed) + the per-record source addresses
  static size_t stage_data(uint32_t recs_per_tile, uint32_t rec_size) {
    return align_up((uint64_t)recs_per_tile * rec_size + (uint64_t)ER_MAX_RUNS * 32 + 64, 128);
  }
  static size_t stage_bytes(uint32_t recs_per_tile, uint32_t rec_size) { return stage_data(recs_per_tile, rec_size) + FE_MAX_RECS * 4; }
  static size_t total(uint32_t recs_per_tile, uint32_t rec_size) {
    return BARS + TABS + ER_STAGES * stage_bytes(recs_per_tile, rec_size) + ER_PW * 3 * ER_MAX_RUNS * 4 +
           2 * (size_t)ER_BATCH * FE_THREADS * 4 + 2 * (size_t)ER_BATCH * sizeof(FoldMeta);
  }
};

__global__ void __launch_bounds__(ER_THREADS, TEZGPU_EMIT_RUNS_MIN_CTAS) k_emit_runs(FastEmitParams fp, uint32_t stage_data, uint32_t stage_bytes) {
  extern __shared__ __align__(128) uint8_t smem_r[];
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_r);
  uint32_t *s_tab = reinterpret_cast<uint32_t *>(smem_r + EmitRunsLayout::BARS);
  uint32_t *s_adv128 = s_tab + 256;
  uint8_t *s_ring = smem_r + EmitRunsLayout::BARS + EmitRunsLayout::TABS;
  uint32_t *s_plan = reinterpret_cast<uint32_t *>(s_ring + (size_t)ER_STAGES * stage_bytes);        // [ER_PW][3][ER_MAX_RUNS]: min, max, base
  uint32_t(*s_part)[FE_THREADS] = reinterpret_cast<uint32_t(*)[FE_THREADS]>(s_plan + ER_PW * 3 * ER_MAX_RUNS);
  FoldMeta *s_meta = reinterpret_cast<FoldMeta *>(reinterpret_cast<uint8_t *>(s_part) + 2 * (size_t)ER_BATCH * FE_THREADS * 4);

  const EmitParams &e = fp.e;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t G = gridDim.x, ntiles = fp.ntiles;
  const uint32_t full0 = (uint32_t)__cvta_generic_to_shared(bars), empty0 = full0 + 8 * ER_STAGES;
  const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(s_ring);
  if (threadIdx.x == 0) {
    for (int s = 0; s < ER_STAGES; s++) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, ER_CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += ER_THREADS) s_tab[i] = e.crc->slice[0][i];
  for (int i = threadIdx.x; i < 4 * 256; i += ER_THREADS) s_adv128[i] = (&e.crc->adv128[0][0])[i];
  __syncthreads();
  if (blockIdx.x >= ntiles) return;
  const TileDesc *__restrict__ tiles = fp.tiles;
  const uint32_t rec_size = e.rec_size;
  const RunTable &rt = e.rec.runs;

  if (warp < ER_PW) {
    // ================================================================ producers: tiles it = warp, warp + ER_PW, ...
    const uint8_t *__restrict__ kv = e.rec.kv;
    const uint64_t kv_safe_end = e.rec.kv_bytes & ~(uint64_t)15;     // bulk copies never read past this offset
    uint32_t *pl_min = s_plan + warp * 3 * ER_MAX_RUNS, *pl_max = pl_min + ER_MAX_RUNS, *pl_base = pl_max + ER_MAX_RUNS;
    uint32_t it = (uint32_t)warp;
    for (uint32_t tile = blockIdx.x + (uint32_t)warp * G; tile < ntiles; tile += ER_PW * G, it += ER_PW) {
      const uint32_t s = it % ER_STAGES, ph = (it / ER_STAGES) & 1u;
      const TileDesc td = tiles[tile];
      const uint32_t seg0 = __ldg(rt.part_seg0 + td.p), nruns = __ldg(rt.part_seg0 + td.p + 1) - seg0;
      // ---- record indices and their runs (8 records per lane)
      uint32_t idx[8], run[8];
      pl_min[lane] = 0xFFFFFFFFu;
      pl_max[lane] = 0u;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const uint32_t j = (uint32_t)lane + 32u * q;
        idx[q] = j < td.nr ? __ldg(e.order + td.r0 + j) : 0u;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const uint32_t j = (uint32_t)lane + 32u * q;
        run[q] = 0;
        if (j < td.nr) {
          uint32_t g = 0;
          while (g + 1 < nruns && idx[q] >= __ldg(rt.rec_base + seg0 + g + 1)) g++;
          run[q] = g;
          atomicMin(&pl_min[g], idx[q]);
          atomicMax(&pl_max[g], idx[q]);
        }
      }
      __syncwarp();
      // ---- lane g plans run g: source byte range, its 16-byte aligned superset, place in the stage
      uint32_t bytes16 = 0, head = 0;
      uint64_t a0 = 0, a1 = 0;
      const bool used = (uint32_t)lane < nruns && pl_min[lane] != 0xFFFFFFFFu;
      if (used) {
        const uint32_t sg = seg0 + (uint32_t)lane;
        const uint64_t so = __ldg(rt.seg_off + sg);
        const uint32_t rb = __ldg(rt.rec_base + sg);
        a0 = so + (uint64_t)(pl_min[lane] - rb) * rec_size;
        a1 = so + (uint64_t)(pl_max[lane] + 1u - rb) * rec_size;
        head = (uint32_t)(a0 & 15u);
        bytes16 = (uint32_t)(((a1 + 15u) & ~(uint64_t)15) - (a0 - head));
      }
      uint32_t incl = bytes16;     // packed placement: exclusive prefix over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t place = incl - bytes16;
      const uint32_t total16 = __shfl_sync(0xffffffffu, incl, 31);
      if (it >= ER_STAGES) mbar_wait(empty0 + 8 * s, ph ^ 1u);     // the consumers are done with this stage
      const uint32_t stg = ring0 + s * stage_bytes;
      pl_base[lane] = stg + place + head;                           // shared address of the run's first needed record
      // bytes the bulk copy may not fetch (they would lie past the end of the input buffer): moved by hand
      uint32_t bulk = bytes16, tail = 0;
      if (used) {
        const uint64_t src0 = a0 - head;
        if (src0 + bytes16 > kv_safe_end) {
          bulk = src0 < kv_safe_end ? (uint32_t)(kv_safe_end - src0) : 0u;
          tail = (uint32_t)(a1 - (src0 + bulk));
          for (uint32_t b = 0; b < tail; b++) sts_b8(stg + place + bulk + b, kv[src0 + bulk + b]);
        }
      }
      uint32_t bulk_total = bulk;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bulk_total += __shfl_xor_sync(0xffffffffu, bulk_total, o);
      __syncwarp();
      // ---- per-record source addresses for the consumers
      uint32_t *s_src = reinterpret_cast<uint32_t *>(s_ring + (size_t)s * stage_bytes + stage_data);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const uint32_t j = (uint32_t)lane + 32u * q;
        if (j < td.nr) s_src[j] = pl_base[run[q]] + (idx[q] - pl_min[run[q]]) * rec_size;
      }
      __syncwarp();
      const uint32_t bar = full0 + 8 * s;
      if (lane == 0) mbar_arrive_expect_tx(bar, bulk_total);        // releases the stores above to the waiting consumers
      __syncwarp();
      if (used && bulk) bulk_copy_g2s(stg + place, kv + (a0 - head), bulk, bar);
      (void)total16;
    }
    return;
  }

  // ================================================================== consumers (256 threads)
  const int tid = threadIdx.x - 32 * ER_PW, cwarp = tid >> 5;
  const uint32_t magic = (uint32_t)((1ull << 32) / rec_size) + 1u;
  CrcChunkFoldT<true> cf;  // the chunk fold's linear maps as warp-resident digit tables (crc32.cuh)
  cf.init(e.crc, lane);
  const uint32_t lane_pow = cf.lane_pow;
  auto consumer_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(FE_THREADS) : "memory"); };
  const DevSmem sm;

  uint32_t it = 0, slot = 0, batch = 0;
  TileDesc tdn = tiles[blockIdx.x];
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += G, it++) {
    const uint32_t s = it % ER_STAGES, ph = (it / ER_STAGES) & 1u;
    const TileDesc td = tdn;
    const bool has1 = tile + G < ntiles;
    if (has1) tdn = tiles[tile + G];
    RunsTileGeom g;
    g.src = reinterpret_cast<const uint32_t *>(s_ring + (size_t)s * stage_bytes + stage_data);
    g.nr = td.nr;
    g.first = td.flags & 1u;
    g.last = td.flags & 2u;
    const uint32_t lead = (uint32_t)(td.abs0 & 15u);
    g.rec0 = lead + (g.first ? 4u : 0u);
    g.body = td.nr * rec_size;
    const uint32_t body_end = g.rec0 + g.body + (g.last ? 2u : 0u);
    const uint32_t cb0 = g.rec0, cb1 = body_end;
    const uint32_t ca = cb0 >> 4, cz = cb1 >> 4;
    uint8_t *dstg = e.out + (td.abs0 - lead);

    mbar_wait(full0 + 8 * s, ph);   // the runs' bytes and the source table have landed

    uint32_t c = 0;
    if (cz > ca) {
      const uint32_t Cn = cz - ca;
      const uint32_t iters = (Cn + FE_THREADS - 1) / FE_THREADS;
      int32_t i = (int32_t)Cn + tid - (int32_t)(iters * FE_THREADS);
      uint8_t *gp = dstg + 16ll * ((int64_t)ca + i);
      for (uint32_t itc = 0; itc < iters; itc++, i += FE_THREADS, gp += 16 * FE_THREADS) {
        if (i + (31 - lane) < 0) continue;  // no lane of this warp owns a chunk yet (first, ragged round only)
        uint4 w = make_uint4(0, 0, 0, 0);
        if (i >= 0) {
          w = runs_assemble(sm, rec_size, magic, g, 16u * (ca + (uint32_t)i));
          if (i == 0) {
            const uint32_t b0 = 16u * ca;
            if (b0 >= lead) stg_stream_v4(gp, w);
            else {
              const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
              for (uint32_t x = lead; x < b0 + 16u; x++) dstg[x] = (uint8_t)(ww[(x - b0) >> 2] >> (8u * ((x - b0) & 3u)));
            }
            const uint32_t skip = cb0 & 15u;
            if (skip) {
              uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (uint32_t q = 0; q < 4; q++) {
                if (skip >= 4 * q + 4) ww[q] = 0;
                else if (skip > 4 * q) ww[q] &= 0xFFFFFFFFu << (8u * (skip - 4 * q));
              }
              w = make_uint4(ww[0], ww[1], ww[2], ww[3]);
            }
          } else {
            stg_stream_v4(gp, w);
          }
        }
        c = cf.fold(c, w, itc + 1 == iters);
      }
    }
    const uint32_t row = (batch & 1u) * ER_BATCH + slot;
    s_part[row][tid] = c;
    if (tid == 0) {
      const uint4 tail = runs_assemble(sm, rec_size, magic, g, 16u * cz);
      const uint32_t tw[4] = {tail.x, tail.y, tail.z, tail.w};
      for (uint32_t x = max(lead, 16u * cz); x < body_end; x++) dstg[x] = (uint8_t)(tw[(x & 15u) >> 2] >> (8u * (x & 3u)));
      if (ca > (lead >> 4))
        for (uint32_t x = lead; x < 16u * ca; x++) dstg[x] = (uint8_t)runs_image_byte(sm, rec_size, magic, g, (int32_t)x - (int32_t)g.rec0);
      FoldMeta m;
      m.tail = tail;
      m.tile = tile;
      m.tiny = cz > ca ? 0u : 1u;
      m.start = cz > ca ? 0u : (cb0 & 15u);
      m.end = cb1 & 15u;
      s_meta[row] = m;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + 8 * s);
    slot++;

    if (slot == (uint32_t)ER_BATCH || !has1) {
      consumer_sync();
      if ((uint32_t)cwarp < slot) {
        const uint32_t r2 = (batch & 1u) * ER_BATCH + (uint32_t)cwarp;
        uint32_t q = 0;
#pragma unroll
        for (int kk = 0; kk < FE_THREADS / 32; kk++) {
          q = s_adv128[q & 0xFF] ^ s_adv128[256 + ((q >> 8) & 0xFF)] ^ s_adv128[512 + ((q >> 16) & 0xFF)] ^ s_adv128[768 + (q >> 24)];
          q ^= s_part[r2][lane + 32 * kk];
        }
        q = crc_multmodp(q, lane_pow);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q ^= __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) {
          const FoldMeta m = s_meta[r2];
          const uint32_t tw[4] = {m.tail.x, m.tail.y, m.tail.z, m.tail.w};
          uint32_t raw = m.tiny ? 0u : q;
          for (uint32_t b = m.start; b < m.end; b++) {
            const uint32_t byte = (tw[b >> 2] >> (8u * (b & 3u))) & 0xFFu;
            raw = s_tab[(raw ^ byte) & 0xFF] ^ (raw >> 8);
          }
          const TileDesc t2 = tiles[m.tile];
          TileCrc tc;
          tc.raw = raw;
          tc.p = t2.p;
          tc.after = t2.after;
          fp.tile_crc[m.tile] = tc;
        }
      }
      slot = 0;
      batch++;
    }
  }
}

static inline bool emit_runs_fits(uint32_t recs_per_tile, uint32_t rec_size, uint32_t max_runs) {
  return max_runs >= 1 && max_runs <= (uint32_t)ER_MAX_RUNS && recs_per_tile <= (uint32_t)FE_MAX_RECS &&
         EmitRunsLayout::total(recs_per_tile, rec_size) <= 110 * 1024;
}

}  // namespace tezgpu
