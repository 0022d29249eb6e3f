// svb200.cu — C ABI (include/svb200.h) and substep driver of the B200 MPM back end.
//
// Mirrors the reference's back-end boundary: CpuState::from_io_state / produce_next_state /
// to_io_state (/root/reference/rust/crates/cpu/src/cpu_state.rs:26-197) with the phase order of
// cpu/src/phase/mod.rs:27-41.  There is NO CPU fallback: every per-particle / per-node phase is a
// CUDA kernel from svb_kernels.cuh; the host only sequences launches, keeps the f64 clock and the
// adaptive time-step history, and builds the (tiny) collider topology / BVH per keyframe.
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/svb200.h"
#include "svb_host.h"
#include "svb_kernels.cuh"

using namespace svb;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t want, bool keep = false) {
    if (want <= bytes) return cudaSuccess;
    const size_t grow = want + want / 4 + 256;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, grow);
    if (e != cudaSuccess) return e;
    if (keep && p && bytes) cudaMemcpy(q, p, bytes, cudaMemcpyDeviceToDevice);
    if (p) cudaFree(p);
    p = q;
    bytes = grow;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

enum Stage : int { ST_MESH = 0, ST_BIN, ST_OFFSETS, ST_PERMUTE, ST_LIMIT, ST_P2G, ST_G2P, ST_ADVANCE, ST_HALO, ST_MIGRATE, ST_COUNT };
const char* const kStageNames[ST_COUNT] = {"mesh_interpolate", "collide_force_bin", "offsets_halo", "invert_zero", "limit_time_step", "p2g", "g2p", "advance", "halo_exchange", "migrate"};

}  // namespace

struct SvbMulti;
struct SvbHandle {
  SvbMulti* multi = nullptr;   // set on the front handle of svb_create_multi: every call fans out to the per-device slab ranks
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_b = nullptr;      // slab ranks: the exchange senders run here, next to P2G / G2P on `stream`
  cudaEvent_t ev_b_start = nullptr, ev_b_end = nullptr;
  SvbConsts consts{};
  SimConsts K{};
  uint32_t n = 0;
  size_t cap = 0;
  DevBuf pbuf[2];
  int cur = 0;
  DevBuf energy;
  std::vector<float> initial_positions;  // never read by the path; echoed by svb_download
  // binning scratch + the two alternating "front sets" of per-substep tile bookkeeping (svb_kernels.cuh: reset_set): substep n runs on
  // set n % 2 while its G2P already bins the advanced positions into the other one
  DevBuf pcell, prank, src_of, grid, melded, node_mask, node_offset, scratch, work_list;
  struct FrontBufs {
    DevBuf table_slots, tile_key, tile_slot, tile_touch, cell_count, slot_first, tile_start, nbr;
    bool fresh = true;      // (re)allocated: memset once before its next use, later uses undo only what they left
  } fs[2];
  size_t tile_cap = 0;      // tiles the per-tile arrays can hold
  int s_cur = 0;            // front set and half of the scalars double buffer of the substep being queued / last completed
  bool binned_ahead = false;  // the OTHER set holds the bins of the current particle state (filled by the last G2P / advance)
  bool limits_ahead = false;  // DtState::next_* hold LimitTimeStepBeforeForce's reductions of the current particle state
  uint32_t table_mask = 0;  // open-addressing slots - 1
  DevBuf dt_state;          // DtState: adaptive time stepping on the device
  DtState* h_dt = nullptr;  // pinned lagged copy
  DevBuf scalars, layer_slots, layer_list;
  StepScalars* h_scalars = nullptr;  // pinned: where the current substep's front-half scalars land (a slot of `lag`)
  cudaEvent_t ev_front = nullptr;
  // peer-memory slab ranks run up to two substeps ahead of the host's look at the front-half scalars, so that a late host thread
  // (eight processes share the box's cores) never leaves its GPU — and through the exchange waits its neighbours — idle
  struct FrontLag { StepScalars* host = nullptr; cudaEvent_t ev = nullptr; double time_before = 0; uint64_t substeps_before = 0; bool was_ahead = false; uint32_t seq_before = 0, dtx_before = 0; } lag[4];
  uint64_t lag_issued = 0, lag_done = 0;
  uint32_t n_ptiles = 0, n_live = 0, n_tiles = 0;
  bool have_grid = false;
  bool store_grid = false, masks_valid = false;
  bool murmur_hash = false;   // option "murmur_table_hash"

  // collider input
  svbh::HostTopology topo;
  bool have_topology = false, have_keyframes = false;
  uint64_t frame = 0;
  float gravity_a[3] = {0, 0, 0}, gravity_b[3] = {0, 0, 0};
  bool has_b = false, has_goals = false;
  DevBuf d_tri, d_opp, d_tri_collider, d_fan_offsets, d_fan_tris;
  DevBuf d_va, d_vb, d_vvel, d_fric_a, d_fric_b, d_damp_a, d_damp_b, d_vpos, d_vnormal, d_tnormal, d_tbox, d_tfric, d_tdamp;
  DevBuf d_node_min, d_node_max, d_node_first, d_node_count, d_children, d_tri_indices;
  DevBuf d_flags_a, d_flags_b, d_goal_a, d_goal_b;
  svbh::FlatBvh bvh;
  MeshDev M{};

  // multi-GPU slabs
  bool slabs = false;
  ncclComm_t comm = nullptr;
  int rank = 0, n_ranks = 1, slab_lo = 0, slab_hi = 0, reach_lo = 0, reach_hi = 0;
  uint64_t orig_offset = 0;
  uint32_t n_global = 0;
  DevBuf comm_counts, halo_send[2], halo_recv[2], mig_send[2], mig_recv[2];
  uint32_t* h_counts = nullptr;  // pinned
  size_t halo_cap = 0, mig_cap = 0, halo_margin = 2048;
  uint64_t halo_tiles_sent = 0, migrated_out = 0;
  // peer-memory exchange (CUDA IPC mailboxes; the NCCL path above stays as the fallback when IPC is unavailable)
  bool p2p = false;
  bool ipc_mapped = false;              // peer_mailbox entries are CUDA-IPC mappings (closed on destroy); false: plain peer pointers of this process
  DevBuf mig_list;                      // slots k_g2p found leaving the slab (left list, right list)
  DevBuf mailbox;                       // this rank's mailbox: SlabHeader | halo in (left, right) | rows in (left, right)
  void* peer_mailbox[16] = {};          // every rank's mailbox mapped into this process (null for self)
  size_t mb_halo_cap = 0, mb_mig_cap = 0, mb_halo_off[2] = {0, 0}, mb_mig_off[2] = {0, 0};
  uint32_t slab_seq = 0;                // message sequence number = slab substeps started
  uint32_t dt_exchanges = 0;            // adaptive time steps: limit exchanges started (the same on every rank)
  uint32_t* n_dev = nullptr;            // device word: rows currently in the particle buffer
  uint32_t* p2p_local = nullptr;        // device scratch of the sending kernels (slot counters, blocks done)

  double time = 0;
  double time_before_last = 0;      // clock before the most recent substep (taken back when that substep turns out to have failed)
  bool failed_rolled_back = false;
  bool device_clock = false;        // this svb_advance call runs adaptive steps: the clock lives in DtState
  uint32_t stop_bits = 0;           // sticky word (errors | ST_STOP_*) that ended the last substep loop
  svbh::AdaptiveTimeStep adaptive;
  uint64_t substeps = 0;
  uint32_t status = 0;
  uint64_t launches = 0;
  std::string last_error;

  // snapshot
  DevBuf snap_p, snap_e;
  double snap_time = 0;
  svbh::AdaptiveTimeStep snap_adaptive;
  uint64_t snap_substeps = 0;
  uint32_t snap_n = 0;
  bool have_snapshot = false;

  // stage timing
  bool timing = false;
  cudaEvent_t ev[ST_COUNT + 1] = {};
  float stage_ms[ST_COUNT] = {};
  cudaEvent_t ev_adv[2] = {};
  float last_advance_ms = 0;
  int last_stage = -1;

  ParticleBuf P(int which) const { return ParticleBuf{pbuf[which].as<uint32_t>(), cap}; }
  ParticleBuf Pc() const { return P(cur); }
};

namespace {

int fail(SvbHandle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->last_error = buf;
  return code;
}
#define CK(call)                                                                                           \
  do {                                                                                                     \
    cudaError_t e_ = (call);                                                                               \
    if (e_ != cudaSuccess) return fail(h, SVB_CUDA_ERROR, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define LAUNCH_CHECK()                                                                                     \
  do {                                                                                                     \
    ++h->launches;                                                                                         \
    cudaError_t e_ = cudaGetLastError();                                                                   \
    if (e_ != cudaSuccess) return fail(h, SVB_CUDA_ERROR, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

inline uint32_t blocks_for(uint64_t n, uint32_t per) { return (uint32_t)((n + per - 1) / per); }

// Launch `kernel` as a programmatic dependent of the kernel queued just before it on `s` (the kernel calls grid_dependency_wait()
// before it touches that kernel's results): its blocks may be scheduled while the predecessor's last blocks are still running, so
// the launch ramp overlaps the predecessor's tail.  SVB_PDL=0 in the environment turns it into an ordinary launch.
template <class... KArgs, class... Args>
cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  static const bool enabled = [] { const char* e = std::getenv("SVB_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

void stage_begin(SvbHandle* h, int st) {
  if (!h->timing) return;
  cudaEventRecord(h->ev[0], h->stream);
  h->last_stage = st;
}
void stage_end(SvbHandle* h) {
  if (!h->timing || h->last_stage < 0) return;
  cudaEventRecord(h->ev[1], h->stream);
  cudaEventSynchronize(h->ev[1]);
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
  h->stage_ms[h->last_stage] += ms;
  h->last_stage = -1;
}

int upload_array(SvbHandle* h, DevBuf& b, const void* src, size_t bytes) {
  CK(b.ensure(bytes ? bytes : 4));
  if (bytes) CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// tile capacity: per-tile arrays + an open-addressing table at <= 25 % load, for both front sets
int ensure_tile_capacity(SvbHandle* h, size_t tiles) {
  if (tiles <= h->tile_cap) return 0;
  const size_t c = tiles + tiles / 2 + 1024;
  size_t slots = 1024;
  while (slots < 4 * c) slots <<= 1;
  for (auto& f : h->fs) {
    CK(f.table_slots.ensure(slots * 16));
    CK(f.tile_key.ensure(c * 8));
    CK(f.tile_slot.ensure(c * 4));
    CK(f.tile_touch.ensure(slots * 4));          // the per-substep particle counters are indexed by table slot (binning never waits for a tile id)
    CK(f.cell_count.ensure(slots * 64 * 4));
    CK(f.slot_first.ensure(slots * 4));
    CK(f.tile_start.ensure((c + 1) * 8));
    CK(f.nbr.ensure(c * 8 * 4));
    f.fresh = true;  // new allocations: the next use memsets them instead of undoing the previous one
  }
  CK(h->grid.ensure(c * 64 * 16));
  CK(h->work_list.ensure((size_t)WORK_CLASSES * c * 4));
  CK(h->node_mask.ensure(c * 8));
  CK(h->node_offset.ensure((c + 1) * 4));
  h->tile_cap = c;
  h->table_mask = (uint32_t)(slots - 1);
  h->binned_ahead = false;   // whatever was binned ahead went with the old tables
  return 0;
}

int set_device(SvbHandle* h) {
  CK(cudaSetDevice(h->device));
  return 0;
}

// CUDA loads a kernel's code lazily, at its first launch, and that load synchronises with the kernels that are running — a deadlock
// when the running kernel is an exchange sender spinning on a counter that only the kernel being loaded will tick (observed: the
// first G2P of a slab run never started while the migration sender waited for its boundary tiles, until the sender's 2 s timeout).
// Every kernel that can be launched next to a waiting one is therefore loaded up front.
int preload_kernels(SvbHandle* h) {
  cudaFuncAttributes a;
#define SVB_PRELOAD(fn) CK(cudaFuncGetAttributes(&a, fn))
  SVB_PRELOAD(k_begin); SVB_PRELOAD(k_bin<false>); SVB_PRELOAD(k_bin<true>); SVB_PRELOAD(k_offsets); SVB_PRELOAD(k_invert_zero<1>); SVB_PRELOAD(k_invert_zero<4>); SVB_PRELOAD(k_touch_nodes);
  SVB_PRELOAD(k_collide_query); SVB_PRELOAD(k_meld); SVB_PRELOAD(k_mesh_lerp); SVB_PRELOAD(k_mesh_tri_normals); SVB_PRELOAD(k_mesh_vertex_normals);
  SVB_PRELOAD((k_collide_cand<1, 1, false>)); SVB_PRELOAD((k_collide_cand<1, 4, true>)); SVB_PRELOAD((k_collide_cand<2, 1, false>)); SVB_PRELOAD((k_collide_cand<2, 4, true>));
  SVB_PRELOAD((k_collide_cand<4, 1, false>)); SVB_PRELOAD((k_collide_cand<4, 4, true>)); SVB_PRELOAD((k_collide_cand<16, 1, false>)); SVB_PRELOAD((k_collide_cand<16, 4, true>));
  SVB_PRELOAD(k_p2g<false>); SVB_PRELOAD(k_p2g<true>);
  SVB_PRELOAD((k_g2p<true, false, true, true, false>)); SVB_PRELOAD((k_g2p<true, false, false, true, true>)); SVB_PRELOAD((k_g2p<true, false, false, true, false>));
  SVB_PRELOAD((k_g2p<true, false, true, false, false>)); SVB_PRELOAD((k_g2p<true, false, false, false, true>)); SVB_PRELOAD((k_g2p<true, false, false, false, false>));
  SVB_PRELOAD((k_g2p<false, true, true, false, false>)); SVB_PRELOAD((k_g2p<false, true, false, false, false>));
  SVB_PRELOAD(k_advance<true>); SVB_PRELOAD(k_advance<false>); SVB_PRELOAD(k_limit_force); SVB_PRELOAD(k_dt_open); SVB_PRELOAD(k_dt_integrate); SVB_PRELOAD(k_dt_tail);
  SVB_PRELOAD(k_halo_send2); SVB_PRELOAD(k_halo_recv2); SVB_PRELOAD(k_migrate_send_list); SVB_PRELOAD(k_migrate_recv); SVB_PRELOAD(k_note_outside); SVB_PRELOAD(k_column_histogram);
#undef SVB_PRELOAD
  return 0;
}

// everything queued for this handle, the exchange senders on the second stream included
cudaError_t sync_streams(SvbHandle* h) {
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess && h->stream_b) e = cudaStreamSynchronize(h->stream_b);
  return e;
}
StepScalars* scalars_of(SvbHandle* h, int which) { return h->scalars.as<StepScalars>() + which; }
StepScalars* cur_scalars(SvbHandle* h) { return scalars_of(h, h->s_cur); }

TileTable tile_table(SvbHandle* h, int which) {
  auto& f = h->fs[which];
  return TileTable{f.table_slots.as<ulonglong2>(), h->table_mask, f.tile_key.as<unsigned long long>(), f.tile_slot.as<uint32_t>(), (uint32_t)h->tile_cap, h->murmur_hash ? 1u : 0u};
}
TileTable tile_table(SvbHandle* h) { return tile_table(h, h->s_cur); }
BinArrays bin_arrays(SvbHandle* h, int which) {
  auto& f = h->fs[which];
  return BinArrays{h->pcell.as<uint32_t>(), f.cell_count.as<uint32_t>(), f.tile_touch.as<uint32_t>(), h->layer_slots.as<unsigned long long>(), h->layer_list.as<uint32_t>()};
}
int memset_fresh_set(SvbHandle* h, int which) {
  auto& f = h->fs[which];
  if (!f.fresh) return 0;
  cudaStream_t s = h->stream;
  CK(cudaMemsetAsync(f.table_slots.p, 0xff, ((size_t)h->table_mask + 1) * 16, s));
  CK(cudaMemsetAsync(f.cell_count.p, 0, ((size_t)h->table_mask + 1) * 64 * 4, s));
  CK(cudaMemsetAsync(f.tile_touch.p, 0, ((size_t)h->table_mask + 1) * 4, s));
  return 0;
}
MeldInfo meld_info(SvbHandle* h) { return MeldInfo{h->layer_slots.as<unsigned long long>(), h->layer_list.as<uint32_t>()}; }

struct StepInputs {
  float factor_b, g[3];
  bool has_mesh;
  bool adaptive;      // the clock, dt, gravity and the frame factor come from DtState on the device
  float dt;           // fixed dt
};
const DtState* dt_ref(SvbHandle* h, const StepInputs& in) { return in.adaptive ? h->dt_state.as<DtState>() : nullptr; }

// Front half of a substep.  Either the set of this substep was filled ahead of time by the previous substep's G2P / advance
// (`binned_ahead`: the substep starts at k_offsets), or it bins now: k_begin, (collide,) k_bin.  Then the cell / tile offsets and the
// halo tiles.  Ends with an async copy of the scalars to pinned memory and an event, so the host can look at the tile count while
// the back half is already queued behind it.  `redo`: the binning of the same substep again after a tile-capacity overflow.
int enqueue_front(SvbHandle* h, const StepInputs& in, bool redo) {
  cudaStream_t s = h->stream;
  const uint32_t n = h->p2p ? (uint32_t)h->cap : h->n;   // peer-memory slabs: the exact row count lives on the device, launch for the capacity
  if (!redo) h->s_cur ^= 1;
  StepScalars* S = cur_scalars(h);
  const StepScalars* S_prev = scalars_of(h, h->s_cur ^ 1);
  const bool ahead = !redo && h->binned_ahead;
  h->binned_ahead = false;
  auto& F = h->fs[h->s_cur];
  const TileTable T = tile_table(h);
  stage_begin(h, ST_BIN);
  if (!ahead) {
    if (F.fresh) {
      if (int rc = memset_fresh_set(h, h->s_cur)) return rc;
      CK(cudaMemsetAsync(h->layer_slots.p, 0, LAYER_SLOTS * 8, s));
    }
    k_begin<<<148, 256, 0, s>>>(S_prev, S, T, F.cell_count.as<uint32_t>(), F.tile_touch.as<uint32_t>(), h->layer_slots.as<unsigned long long>(), n, h->p2p ? h->n_dev : nullptr, F.fresh ? 1 : 0);
    LAUNCH_CHECK();
    F.fresh = false;
    const uint32_t blocks = std::max<uint32_t>(blocks_for(n, 256), 1);
    if (in.has_mesh && !redo) {  // collide.rs:21-206 (once per substep: it edits v and the bits in place)
      uint32_t* candidates = h->prank.as<uint32_t>();   // free until k_bin writes the ranks
      k_collide_query<<<blocks, 256, 0, s>>>(h->Pc(), S, h->K, h->M, candidates, (uint32_t)h->cap, n);
      LAUNCH_CHECK();
      const uint32_t nc = h->M.n_colliders;
#define SVB_CAND(NC)                                                                                                                    \
  do {                                                                                                                                \
    k_collide_cand<NC, 1, false><<<148 * 8, 128, 0, s>>>(h->Pc(), S, h->K, h->M, candidates, (uint32_t)h->cap, in.dt, dt_ref(h, in)); \
    k_collide_cand<NC, 4, true><<<148 * 16, 128, 0, s>>>(h->Pc(), S, h->K, h->M, candidates, (uint32_t)h->cap, in.dt, dt_ref(h, in)); \
    ++h->launches;                                                                                                                    \
  } while (0)
      if (nc <= 1) SVB_CAND(1); else if (nc <= 2) SVB_CAND(2); else if (nc <= 4) SVB_CAND(4); else SVB_CAND(16);
#undef SVB_CAND
      LAUNCH_CHECK();
    }
    const BinArrays B = bin_arrays(h, h->s_cur);
    if (in.has_mesh) k_bin<true><<<blocks, 256, 0, s>>>(h->Pc(), S, h->K, T, B, n);
    else k_bin<false><<<blocks, 256, 0, s>>>(h->Pc(), S, h->K, T, B, n);
    LAUNCH_CHECK();
  }
  stage_end(h);
  stage_begin(h, ST_OFFSETS);
  const uint32_t lag = std::max<uint32_t>(h->n_ptiles, 1024);
  static const bool by_size = [] { const char* e = std::getenv("SVB_WORK_ORDER"); return !(e && e[0] == '0'); }();
  const SlabColumns cols{h->slab_lo, h->slab_hi, h->work_list.as<uint32_t>(), h->p2p ? 1 : 0, by_size ? 1 : 0};
  const uint32_t offsets_grid = std::min<uint32_t>(blocks_for((uint64_t)lag * 32 * 2, 256), 148 * 8);
  if (ahead && !h->slabs && !h->timing)   // straight behind the previous substep's G2P: a programmatic dependent of it
    CK(launch_dependent(k_offsets, dim3(offsets_grid), dim3(256), 0, s, S, S_prev, n, (const uint32_t*)nullptr, T, F.cell_count.as<uint32_t>(), F.tile_start.as<uint2>(), F.slot_first.as<uint32_t>(),
                        (const uint32_t*)F.tile_touch.as<uint32_t>(), F.nbr.as<int>(), cols));
  else
    k_offsets<<<offsets_grid, 256, 0, s>>>(S, ahead ? S_prev : nullptr, n, h->p2p ? h->n_dev : nullptr, T, F.cell_count.as<uint32_t>(), F.tile_start.as<uint2>(), F.slot_first.as<uint32_t>(),
                                           F.tile_touch.as<uint32_t>(), F.nbr.as<int>(), cols);
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(h->h_scalars, S, sizeof(StepScalars), cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(h->ev_front, s));
  stage_end(h);
  return 0;
}

// re-bin + grid preparation that follows the front half; `prepare_next`: also reset the other front set, which this substep's
// G2P / advance will fill with the bins of the advanced positions
int enqueue_rebin(SvbHandle* h, bool prepare_next) {
  cudaStream_t s = h->stream;
  const uint32_t n = h->p2p ? (uint32_t)h->cap : h->n;
  StepScalars* S = cur_scalars(h);
  auto& F = h->fs[h->s_cur];
  stage_begin(h, ST_PERMUTE);
  PrepareNext prep{};
  if (prepare_next) {
    const int nx = h->s_cur ^ 1;
    auto& N = h->fs[nx];
    if (int rc = memset_fresh_set(h, nx)) return rc;
    prep = PrepareNext{scalars_of(h, nx), tile_table(h, nx), N.cell_count.as<uint32_t>(), N.tile_touch.as<uint32_t>(), n, h->p2p ? h->n_dev : nullptr, N.fresh ? 1 : 0, 74u};
    N.fresh = false;
  }
  static const int rows_per_thread = [] { const char* e = std::getenv("SVB_INVERT_ROWS"); return e && e[0] == '1' ? 1 : 4; }();
  const uint32_t invert_blocks = blocks_for(n, 256 * rows_per_thread);
  if (rows_per_thread == 4)
    k_invert_zero<4><<<invert_blocks + 148 * 2 + prep.blocks, 256, 0, s>>>(S, h->pcell.as<uint32_t>(), F.cell_count.as<uint32_t>(), F.slot_first.as<uint32_t>(), h->src_of.as<uint32_t>(), n,
                                                                          invert_blocks, h->grid.as<float4>(), h->store_grid ? h->node_mask.as<unsigned long long>() : nullptr, (uint32_t)h->tile_cap, prep);
  else
    k_invert_zero<1><<<invert_blocks + 148 * 2 + prep.blocks, 256, 0, s>>>(S, h->pcell.as<uint32_t>(), F.cell_count.as<uint32_t>(), F.slot_first.as<uint32_t>(), h->src_of.as<uint32_t>(), n,
                                                                          invert_blocks, h->grid.as<float4>(), h->store_grid ? h->node_mask.as<unsigned long long>() : nullptr, (uint32_t)h->tile_cap, prep);
  LAUNCH_CHECK();
  h->masks_valid = false;
  if (h->store_grid) {
    k_touch_nodes<<<148 * 8, 256, 0, s>>>(h->Pc(), h->src_of.as<uint32_t>(), F.tile_start.as<uint2>(), F.nbr.as<int>(), S, h->K.h, h->node_mask.as<unsigned long long>());
    LAUNCH_CHECK();
    h->masks_valid = true;
  }
  stage_end(h);
  return 0;
}

// the tiles a P2G / G2P launch works through: all particle tiles, or (peer-memory slab ranks) the boundary / interior list of k_offsets
// how finely a launch cuts the tiles' particle runs into work items (svb_kernels.cuh: WorkList::parts).  Whole tiles: measured on
// B200 (profiles/README.md r2i) halves and quarters are SLOWER at 1 M and at 8 M particles (G2P 78 -> 88 -> 112 us at 1 M) — every
// work item pays ~3 us of dependent round trips (claim, neighbour ids, velocity tile, row indices) that more items only multiply,
// which outweighs the fuller tail.  SVB_PARTS=2|4 in the environment keeps the experiment reproducible.
uint32_t work_parts(SvbHandle*, uint32_t) {
  static const int forced = [] { const char* e = std::getenv("SVB_PARTS"); return e ? std::atoi(e) : 0; }();
  return forced == 2 || forced == 4 ? (uint32_t)forced : 1u;
}
// ... and how many of the last tiles of a launch are cut that way: SVB_SPLIT_LAST (tiles; "all" = every tile)
uint32_t work_split_last(int phase) {
  static const uint32_t v = [] {
    const char* e = std::getenv("SVB_SPLIT_LAST");
    if (!e) return 0u;
    return e[0] == 'a' ? 0xffffffffu : (uint32_t)std::atoi(e);
  }();
  return v ? v : 148u * (phase == 0 ? P2G_CTAS_PER_SM : G2P_CTAS_PER_SM);   // default: as many tiles as CTAs are resident = the launch's final wave
}
WorkList work_all(SvbHandle* h, int phase, int tail) {
  StepScalars* S = cur_scalars(h);
  return WorkList{h->work_list.as<uint32_t>(), (uint32_t)h->tile_cap, S->n_class, &S->work_counter[phase], nullptr, tail, work_parts(h, phase == 0 ? P2G_CTAS_PER_SM : G2P_CTAS_PER_SM), work_split_last(phase)};
}
// boundary tiles first, then the interior ones; `phase` 0 = P2G, 1 = G2P (work cursor and boundary-done counter of the scalars)
WorkList work_ordered(SvbHandle* h, int phase, int tail) {
  StepScalars* S = cur_scalars(h);
  return WorkList{h->work_list.as<uint32_t>(), (uint32_t)h->tile_cap, S->n_class, &S->work_counter[phase], &S->boundary_done[phase], tail, work_parts(h, phase == 0 ? P2G_CTAS_PER_SM : G2P_CTAS_PER_SM), work_split_last(phase)};
}

int enqueue_p2g(SvbHandle* h, const StepInputs& in, const WorkList& W, uint32_t grid_cap = 148 * P2G_CTAS_PER_SM) {
  cudaStream_t s = h->stream;
  auto& F = h->fs[h->s_cur];
  const uint32_t lag = std::max<uint32_t>(h->n_ptiles, 1);
  const uint32_t p2g_grid = std::max<uint32_t>(148, std::min<uint32_t>(lag * 2 * W.parts, grid_cap));
  ForceIn force{};
  force.dt = in.dt; force.gx = in.g[0]; force.gy = in.g[1]; force.gz = in.g[2]; force.factor_b = in.factor_b;
  force.D = dt_ref(h, in);
  if (h->has_goals) {
    force.G.flags_a = h->d_flags_a.as<uint32_t>();
    force.G.flags_b = h->has_b ? h->d_flags_b.as<uint32_t>() : force.G.flags_a;
    force.G.goal_a = h->d_goal_a.as<float>();
    force.G.goal_b = h->has_b ? h->d_goal_b.as<float>() : force.G.goal_a;
    CK(launch_dependent(k_p2g<true>, dim3(p2g_grid), dim3(P2G_WARPS * 32), P2G_SMEM, s, h->Pc(), h->src_of.as<uint32_t>(), F.tile_start.as<uint2>(), F.nbr.as<int>(), cur_scalars(h), h->grid.as<float4>(), h->K.h,
                        in.dt, force, W));
  } else {
    CK(launch_dependent(k_p2g<false>, dim3(p2g_grid), dim3(P2G_WARPS * 32), P2G_SMEM, s, h->Pc(), h->src_of.as<uint32_t>(), F.tile_start.as<uint2>(), F.nbr.as<int>(), cur_scalars(h), h->grid.as<float4>(), h->K.h,
                        in.dt, force, W));
  }
  LAUNCH_CHECK();
  return 0;
}

// the substep that was just queued turned out to be a no-op on the device (an earlier substep stopped the run): forget it
void forget_noop_substep(SvbHandle* h, bool was_ahead, bool back_enqueued) {
  if (back_enqueued) h->cur ^= 1;   // the queued G2P wrote nothing: undo the buffer swap
  h->s_cur ^= 1;
  h->binned_ahead = was_ahead;      // k_offsets aborted before touching the set: its bins still describe the particle state
}

// The substep that raised a simulation-level error does not count: the reference returns from the failing phase without cycling it
// or advancing the clock (cpu/src/cpu_state.rs:176-190), so the stored frame carries the time BEFORE that substep.  The device
// finishes the failing substep (the error is only seen afterwards), hence the host takes its bookkeeping back, once.
void rollback_failed_substep(SvbHandle* h) {
  if (h->failed_rolled_back || h->substeps == 0) return;
  h->time = h->time_before_last;
  --h->substeps;
  h->failed_rolled_back = true;
}

// wait for the front half's scalars; on tile overflow grow the capacity and redo the binning
// (the state is only rewritten by G2P, and k_invert / P2G / G2P no-op on overflow)
//   0: fine (back half queued)   1: fine, the caller still has to queue the back half   2: the run was stopped by an earlier substep
int settle_front(SvbHandle* h, const StepInputs& in, bool back_enqueued, bool was_ahead) {
  for (int attempt = 0;; ++attempt) {
    CK(cudaEventSynchronize(h->ev_front));
    const StepScalars& r = *h->h_scalars;
    if (r.status & ST_KEY_RANGE) return fail(h, SVB_KEY_RANGE, "a live particle lies outside the +-2^18 grid-cell range of the tile keys (or crossed more than one slab in a substep)");
    if (r.status & (ST_COMM_TIMEOUT | ST_COMM_OVERFLOW)) return fail(h, SVB_COMM_ERROR, "slab exchange failed in an earlier substep (status 0x%x)", r.status);
    if (r.status & ST_ZERO_DT) return fail(h, SVB_ZERO_TIME_STEP, "The time step ended up being 0");
    if (r.sticky) {  // an earlier substep stopped the run: this one was a no-op on the device
      h->status |= (r.sticky | r.accum) & 0xffffu;
      CK(cudaStreamSynchronize(h->stream));
      forget_noop_substep(h, was_ahead, back_enqueued);
      if (r.sticky & 0xffffu) rollback_failed_substep(h);
      h->stop_bits = r.sticky;
      return 2;
    }
    if (!(r.status & ST_TILE_OVERFLOW)) {
      h->n_tiles = r.n_tiles;
      h->n_ptiles = r.n_ptiles;
      h->n_live = r.n_live;
      h->status |= (r.status | r.accum) & 0xffffu;
      return back_enqueued ? 0 : 1;  // 1: caller still has to enqueue the back half
    }
    if (attempt > 8) return fail(h, SVB_CUDA_ERROR, "tile capacity did not settle");
    CK(cudaStreamSynchronize(h->stream));
    if (back_enqueued) h->cur ^= 1;  // the queued G2P was a no-op: undo the buffer swap
    back_enqueued = false;
    if (int rc = ensure_tile_capacity(h, (size_t)r.n_tiles * 2 + 1024)) return rc;
    if (int rc = enqueue_front(h, in, /*redo=*/true)) return rc;
  }
}

// MeldGrid (collider scenes: the sibling layers are melded into a velocity grid first; without colliders G2P divides by the mass
// while it stages a tile)
int enqueue_meld(SvbHandle* h) {
  CK(h->melded.ensure(h->tile_cap * 64 * 16));
  k_meld<<<148 * 8, 256, 0, h->stream>>>(cur_scalars(h), h->grid.as<float4>(), h->melded.as<float4>(), tile_table(h), meld_info(h));
  LAUNCH_CHECK();
  return 0;
}
// CollectVelocity (+ Advance + Cull when fused) over the tiles of `W`, from particle buffer `src_buf` into `dst_buf` (the caller
// swaps `cur` once all launches of the substep are queued).  `bin_next`: the fused kernel also bins the advanced positions into the
// other front set.
int enqueue_g2p(SvbHandle* h, bool has_mesh, bool fuse, float dt, bool bin_next, const WorkList& W, int src_buf, const MigrateCut* cut = nullptr, uint32_t grid_cap = 148 * 16) {
  cudaStream_t s = h->stream;
  StepScalars* S = cur_scalars(h);
  auto& F = h->fs[h->s_cur];
  const float4* src = has_mesh ? h->melded.as<float4>() : h->grid.as<float4>();
  const ParticleBuf P = h->P(src_buf), D = h->P(src_buf ^ 1);
  const uint32_t* src_of = h->src_of.as<uint32_t>();
  const uint2* tile_start = F.tile_start.as<uint2>();
  float* en = h->energy.as<float>();
  const int* nb = F.nbr.as<int>();
  const uint32_t lag = std::max<uint32_t>(h->n_ptiles, 1);
  const uint32_t g2p_grid = std::max<uint32_t>(148, std::min<uint32_t>(lag * 2 * W.parts, grid_cap));
  const MigrateCut mc = cut ? *cut : MigrateCut{};
  BinNext bn{};
  if (bin_next) bn = BinNext{scalars_of(h, h->s_cur ^ 1), tile_table(h, h->s_cur ^ 1), bin_arrays(h, h->s_cur ^ 1)};
#define SVB_G2P(F_, R_, M_, SL_, B_) \
  CK(launch_dependent(k_g2p<F_, R_, M_, SL_, B_>, dim3(g2p_grid), dim3(G2P_THREADS), 0, s, P, D, src_of, en, tile_start, nb, S, src, h->K, dt, mc, bn, F.tile_key.as<unsigned long long>(), W))
  if (fuse && cut) {
    if (has_mesh) SVB_G2P(true, false, true, true, false);
    else if (bin_next) SVB_G2P(true, false, false, true, true);
    else SVB_G2P(true, false, false, true, false);
  } else if (fuse) {
    if (has_mesh) SVB_G2P(true, false, true, false, false);
    else if (bin_next) SVB_G2P(true, false, false, false, true);
    else SVB_G2P(true, false, false, false, false);
  } else {
    if (has_mesh) SVB_G2P(false, true, true, false, false);
    else SVB_G2P(false, true, false, false, false);
  }
#undef SVB_G2P
  LAUNCH_CHECK();
  return 0;
}

int halo_exchange(SvbHandle* h);
int migrate(SvbHandle* h);
// svb_multi.inl: a handle over several devices, its calls fan out to one slab rank per device
int multi_upload(SvbHandle* front, const SvbParticles* p, double time);
int multi_download(SvbHandle* front, SvbParticles* out);
int multi_advance(SvbHandle* front, double target_time, float max_time_step, int32_t adaptive, const volatile int32_t* cancel, void (*progress)(void*, size_t), void* user);
int multi_set_topology(SvbHandle* front, uint32_t n_colliders, const uint32_t* num_vertices, const uint32_t* num_triangles, const uint32_t* triangles);
int multi_set_keyframes(SvbHandle* front, uint64_t frame, const SvbKeyframe* a, const SvbKeyframe* b);
void multi_set_option(SvbHandle* front, const char* name, double value);
void multi_destroy(SvbHandle* front);
int mailbox_alloc(SvbHandle* h, size_t n_max);
int mailbox_finish(SvbHandle* h);
int substep_slab(SvbHandle* h, const StepInputs& in);
int substep_slab_p2p(SvbHandle* h, const StepInputs& in);
int setup_peer_mailboxes(SvbHandle* h);
int resize_particles(SvbHandle* h, size_t new_cap);

// slab ranks exchange the limit reductions of adaptive time stepping through the mailbox headers (k_dt_*: dt_exchange)
DtPeers dt_peers(SvbHandle* h) {
  DtPeers p{};
  p.rank = h->rank;
  p.n_ranks = h->p2p ? h->n_ranks : 1;
  if (h->p2p) {
    p.mine = h->mailbox.as<SlabHeader>();
    for (int r = 0; r < h->n_ranks; ++r)
      if (r != h->rank) {
        SlabHeader* ph = reinterpret_cast<SlabHeader*>(h->peer_mailbox[r]);
        p.post_seq[r] = &ph->dt_seq[0][h->rank];
        p.post_val[r] = ph->dt_val[0][h->rank];
      }
    p.exchange = ++h->dt_exchanges;
  }
  return p;
}

int enqueue_mesh(SvbHandle* h, const StepInputs& in) {
  cudaStream_t s = h->stream;
  stage_begin(h, ST_MESH);
  const uint32_t m = std::max(h->topo.n_vertices * 3, h->topo.n_triangles);
  k_mesh_lerp<<<blocks_for(m, 256), 256, 0, s>>>(h->M, in.factor_b, dt_ref(h, in));
  LAUNCH_CHECK();
  k_mesh_tri_normals<<<blocks_for(h->topo.n_triangles, 256), 256, 0, s>>>(h->M);
  LAUNCH_CHECK();
  k_mesh_vertex_normals<<<blocks_for(h->topo.n_vertices, 256), 256, 0, s>>>(h->M);
  LAUNCH_CHECK();
  stage_end(h);
  return 0;
}

// ---- one substep: the 12 phases of cpu/src/phase/mod.rs:27-41 in the reference's order.
// Returns 0, a negative fatal status, or 2 when the run was stopped by an earlier substep (this one was a no-op).
int substep(SvbHandle* h, bool adaptive_steps) {
  cudaStream_t s = h->stream;
  const uint32_t n = h->n;
  StepInputs in{};
  in.adaptive = adaptive_steps;
  in.has_mesh = h->topo.n_triangles > 0;
  if (!adaptive_steps) {
    if (h->adaptive.allowed() == 0.f) return fail(h, SVB_ZERO_TIME_STEP, "The time step ended up being 0");
    // -- InterpolateInput (interpolate_input.rs:18-107; frame factor xpu/src/frame_input.rs:266-278)
    const double frame_time = h->time * (double)h->consts.frames_per_second;
    const uint64_t frame_low = (uint64_t)std::floor(frame_time);
    if (frame_low != h->frame) return fail(h, SVB_FRAME_INPUT, "Wrong frame loaded: %llu (need %llu)", (unsigned long long)h->frame, (unsigned long long)frame_low);
    in.factor_b = (float)std::fmod(frame_time, 1.0);
    const float factor_a = 1.f - in.factor_b;
    const float* gb = h->has_b ? h->gravity_b : h->gravity_a;
    for (int k = 0; k < 3; ++k) in.g[k] = factor_a * h->gravity_a[k] + in.factor_b * gb[k];
    in.dt = h->adaptive.allowed();
  }
  if (in.has_mesh)
    if (int rc = enqueue_mesh(h, in)) return rc;
  if (h->slabs) {
    if (adaptive_steps && !h->p2p) return fail(h, SVB_BAD_ARGUMENT, "adaptive time steps on slab ranks need the peer-memory exchange path (SVB_SLAB_NCCL forces the NCCL fallback)");
    return h->p2p ? substep_slab_p2p(h, in) : substep_slab(h, in);
  }
  if (n == 0) {
    if (adaptive_steps) return fail(h, SVB_BAD_ARGUMENT, "adaptive time steps need at least one particle");
    h->time_before_last = h->time;
    h->time += (double)h->adaptive.allowed();
    ++h->substeps;
    return 0;
  }

  // -- Sort + UpdateGridNodes (+ Collide): filled ahead by the previous substep's G2P, or binned now.  Collide / force are per
  //    particle, so running them in the pre-bin order is equivalent to the reference's Sort -> Collide -> Force.
  const bool was_ahead = h->binned_ahead;
  const bool bin_next = !in.has_mesh;   // with a mesh the collider bits of the next substep are not known yet
  if (int rc = enqueue_front(h, in, /*redo=*/false)) return rc;
  for (int pass = 0;; ++pass) {
    // the whole back half is queued behind the front half, then the host looks at the front half's result
    if (int rc = enqueue_rebin(h, bin_next)) return rc;
    stage_begin(h, ST_P2G);
    if (int rc = enqueue_p2g(h, in, work_all(h, 0, 0))) return rc;
    stage_end(h);
    stage_begin(h, ST_G2P);
    if (in.has_mesh)
      if (int rc = enqueue_meld(h)) return rc;
    if (int rc = enqueue_g2p(h, in.has_mesh, /*fuse=*/!adaptive_steps, in.dt, bin_next && !adaptive_steps, work_all(h, 1, 1), h->cur)) return rc;
    h->cur ^= 1;  // the binned buffer written by G2P is the current one from here on
    stage_end(h);
    if (adaptive_steps) {
      // -- LimitTimeStepBeforeIntegrate, AdvanceParticles + CullParticles, the clock: all on the device
      DtState* D = h->dt_state.as<DtState>();
      stage_begin(h, ST_LIMIT);
      k_dt_integrate<<<1, 1, 0, s>>>(D, cur_scalars(h), dt_peers(h));
      LAUNCH_CHECK();
      stage_end(h);
      stage_begin(h, ST_ADVANCE);
      BinNext bn{};
      if (bin_next) bn = BinNext{scalars_of(h, h->s_cur ^ 1), tile_table(h, h->s_cur ^ 1), bin_arrays(h, h->s_cur ^ 1)};
      if (bin_next) k_advance<true><<<blocks_for(n, 256), 256, 0, s>>>(h->Pc(), h->energy.as<float>(), cur_scalars(h), h->K, n, D, bn, MigrateCut{});
      else k_advance<false><<<blocks_for(n, 256), 256, 0, s>>>(h->Pc(), h->energy.as<float>(), cur_scalars(h), h->K, n, D, bn, MigrateCut{});
      LAUNCH_CHECK();
      k_dt_tail<<<1, 1, 0, s>>>(D, cur_scalars(h), dt_peers(h));
      LAUNCH_CHECK();
      CK(cudaMemcpyAsync(h->h_dt, D, sizeof(DtState), cudaMemcpyDeviceToHost, s));
      stage_end(h);
    }
    if (pass > 0) break;
    const int rc = settle_front(h, in, /*back_enqueued=*/true, was_ahead);
    if (rc < 0) return rc;
    if (rc == 2) return 2;  // stopped by an earlier substep; time does not advance
    if (rc == 0) break;
    // rc == 1: the binning was redone with a larger tile capacity: queue the back half again
  }
  h->binned_ahead = bin_next;
  h->limits_ahead = adaptive_steps;
  h->have_grid = true;
  if (!adaptive_steps) {
    h->time_before_last = h->time;
    h->time += (double)in.dt;
    ++h->substeps;
  }
  return 0;
}

// one fixed-dt substep of a slab rank: like the single-GPU flow, with the halo exchange between P2G
// and G2P and the particle migration after the advance.  Runs even with zero resident particles
// (the neighbours still expect this rank's messages).
int substep_slab(SvbHandle* h, const StepInputs& in) {
  cudaStream_t s = h->stream;
  const float dt = in.dt;
  bool redo = false;
  for (int attempt = 0;; ++attempt) {
    if (int rc = enqueue_front(h, in, redo)) return rc;
    const int rc = settle_front(h, in, /*back_enqueued=*/false, false);
    if (rc < 0) return rc;
    if (rc == 2) return 2;
    // tiles arriving with the halo need room in the table; growing it loses its contents, so re-bin
    const size_t want = (size_t)h->n_tiles + h->halo_margin;
    if (want <= h->tile_cap) break;
    if (attempt > 4) return fail(h, SVB_COMM_ERROR, "tile capacity did not settle");
    CK(cudaStreamSynchronize(s));
    if (int rc2 = ensure_tile_capacity(h, want + want / 2)) return rc2;
    redo = true;
  }
  const uint32_t n_after = h->h_scalars->n_live + h->h_scalars->n_tomb;  // rows of migrated particles are dropped by the re-bin
  if (int rc = enqueue_rebin(h, false)) return rc;
  stage_begin(h, ST_P2G);
  if (int rc = enqueue_p2g(h, in, work_all(h, 0, 0))) return rc;
  stage_end(h);
  stage_begin(h, ST_HALO);
  if (int rc = halo_exchange(h)) return rc;
  stage_end(h);
  stage_begin(h, ST_G2P);
  if (in.has_mesh)
    if (int rc = enqueue_meld(h)) return rc;
  if (int rc = enqueue_g2p(h, in.has_mesh, /*fuse=*/true, dt, false, work_all(h, 1, 1), h->cur)) return rc;
  h->cur ^= 1;
  h->n = n_after;
  // a FAILED particle on any rank stops every rank after this substep
  uint32_t* flag = h->comm_counts.as<uint32_t>() + 8;
  CK(cudaMemcpyAsync(flag, &cur_scalars(h)->sticky_new, 4, cudaMemcpyDeviceToDevice, s));
  if (ncclAllReduce(flag, flag, 1, ncclUint32, ncclMax, h->comm, s) != ncclSuccess) return fail(h, SVB_COMM_ERROR, "ncclAllReduce failed");
  CK(cudaMemcpyAsync(&cur_scalars(h)->sticky_new, flag, 4, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemcpyAsync(h->h_counts + 8, flag, 4, cudaMemcpyDeviceToHost, s));
  stage_end(h);
  stage_begin(h, ST_MIGRATE);
  if (int rc = migrate(h)) return rc;  // synchronises the stream
  stage_end(h);
  h->status |= h->h_counts[8] & 0xffffu;
  h->have_grid = true;
  h->time_before_last = h->time;
  h->time += (double)dt;
  ++h->substeps;
  return 0;
}

// The host's (lagged) look at the front-half scalars of the oldest unchecked substep of a peer-memory slab rank.
//   0: fine   2: that substep was a no-op (an earlier one stopped the run); every substep queued since is forgotten
int process_front_p2p(SvbHandle* h) {
  SvbHandle::FrontLag& L = h->lag[h->lag_done % 4];
  CK(cudaEventSynchronize(L.ev));
  const StepScalars& r = *L.host;
  if (r.status & ST_KEY_RANGE) return fail(h, SVB_KEY_RANGE, "a live particle lies outside the +-2^18 grid-cell range of the tile keys (or crossed more than one slab in a substep)");
  if (r.status & ST_TILE_OVERFLOW) return fail(h, SVB_COMM_ERROR, "tile capacity exceeded on a slab rank (%u tiles > %zu)", r.n_tiles, h->tile_cap);
  if (r.status & (ST_COMM_TIMEOUT | ST_COMM_OVERFLOW)) {   // raised by an earlier substep's back half and carried forward
    StepScalars both[2];
    uint32_t tr[16] = {};
    sync_streams(h);
    cudaMemcpy(both, h->scalars.p, sizeof both, cudaMemcpyDeviceToHost);
    cudaMemcpyFromSymbol(tr, g_exchange_trace, sizeof tr);
    return fail(h, SVB_COMM_ERROR, "%s (rank %d, message %u; last message numbers seen by halo send in/waited/published %u %u %u, halo recv in/got %u %u, migrate send in/waited/published %u %u %u, migrate recv in/got %u %u; scalars halves: boundary tiles %u / %u, interior %u / %u, ticks P2G %u / %u, G2P %u / %u, status 0x%x / 0x%x)",
                r.status & ST_COMM_TIMEOUT ? "a neighbour slab's message did not arrive" : "a slab mailbox or the particle buffer ran out of room", h->rank, h->slab_seq, tr[0], tr[1], tr[2], tr[3],
                tr[4], tr[5], tr[6], tr[7], tr[8], tr[9], both[0].n_work[0], both[1].n_work[0], both[0].n_work[1], both[1].n_work[1], both[0].boundary_done[0], both[1].boundary_done[0], both[0].boundary_done[1], both[1].boundary_done[1], both[0].status,
                both[1].status);
  }
  if (r.sticky) {  // the previous substep failed somewhere: this one and everything queued behind it were no-ops on every rank
    h->status |= (r.sticky | r.accum) & 0xffffu;
    CK(sync_streams(h));
    const uint64_t noops = h->lag_issued - h->lag_done;
    if (noops & 1) { h->cur ^= 1; h->s_cur ^= 1; }   // each queued G2P swapped the buffers and each front the set: undo (they wrote nothing)
    h->binned_ahead = L.was_ahead;
    h->slab_seq = L.seq_before;       // the no-op substeps sent nothing (the exchange kernels return on a stopped run): every rank
    h->dt_exchanges = L.dtx_before;   // forgets its own — possibly different — number of them and the message numbers stay in step
    // the clock of the substep that raised the error does not count either (cpu_state.rs:176-190)
    const SvbHandle::FrontLag& failing = h->lag_done > 0 ? h->lag[(h->lag_done - 1) % 4] : L;
    if (!h->device_clock) {
      h->time = failing.time_before;
      h->substeps = failing.substeps_before;
    }
    h->failed_rolled_back = true;
    h->stop_bits = r.sticky;
    h->lag_done = h->lag_issued;
    return 2;
  }
  h->n_tiles = r.n_tiles;
  h->n_ptiles = r.n_ptiles;
  h->n_live = r.n_live;
  h->status |= (r.status | r.accum) & 0xffffu;
  ++h->lag_done;
  if (((size_t)r.n_tiles + h->halo_margin) * 3 / 2 > h->tile_cap) {  // grow ahead of need: an overflow cannot be redone once messages are out
    CK(sync_streams(h));   // (every substep queued so far completes first: they are real ones)
    while (h->lag_done < h->lag_issued) {
      const int rc = process_front_p2p(h);
      if (rc) return rc;
    }
    if (int rc = ensure_tile_capacity(h, ((size_t)r.n_tiles + h->halo_margin) * 3)) return rc;   // (drops what was binned ahead: the next substep bins with k_bin)
  }
  return 0;
}

// one substep of a slab rank over peer memory: nothing on the data path returns to the host, and the exchanges hide behind the
// interior tiles' work.  P2G and G2P take the boundary tiles first; the two SENDING kernels run on a second stream next to them,
// wait (on the device) for the last boundary tile and ship the halo columns / the leavers while the interior is still in progress:
//   main stream:   front | rebin | P2G [boundary, interior] | halo receive | (meld) G2P [boundary, interior] | migration receive
//   second stream:         ......  halo send (after the boundary tiles)  ......  migration send (after the boundary tiles)
// so both receives find the neighbour's message already there.  The whole substep is queued, then the host looks at the front half
// of an EARLIER substep (up to two stay in flight).
int substep_slab_p2p(SvbHandle* h, const StepInputs& in) {
  cudaStream_t s = h->stream, sb = h->stream_b;
  const float dt = in.dt;
  SvbHandle::FrontLag& L = h->lag[h->lag_issued % 4];
  L.time_before = h->time; L.substeps_before = h->substeps; L.was_ahead = h->binned_ahead; L.seq_before = h->slab_seq; L.dtx_before = h->dt_exchanges;
  const uint32_t seq = ++h->slab_seq;
  const bool bin_next = !in.has_mesh;
  h->h_scalars = L.host;
  h->ev_front = L.ev;
  CK(cudaStreamWaitEvent(s, h->ev_b_end, 0));   // the previous substep's senders are done with the scalars / buffers this one recycles
  if (int rc = enqueue_front(h, in, /*redo=*/false)) return rc;
  if (int rc = enqueue_rebin(h, bin_next)) return rc;
  CK(cudaEventRecord(h->ev_b_start, s));
  CK(cudaStreamWaitEvent(sb, h->ev_b_start, 0));
  StepScalars* S = cur_scalars(h);
  const TileTable T = tile_table(h);
  SlabHeader* my_hdr = h->mailbox.as<SlabHeader>();
  unsigned char* my_mb = h->mailbox.as<unsigned char>();
  const bool has[2] = {h->rank > 0, h->rank + 1 < h->n_ranks};
  const bool concurrent = !h->timing;   // (the instrumented pass times the stages one after the other on the main stream)
  // ---- second stream: the halo sender, gated on the device by P2G's boundary tiles
  stage_begin(h, ST_P2G);
  if (concurrent && (has[0] || has[1])) {
    HaloPeers hp{};
    for (int side = 0; side < 2; ++side)
      if (has[side]) {
        unsigned char* peer = static_cast<unsigned char*>(h->peer_mailbox[h->rank + (side ? 1 : -1)]);
        SlabHeader* ph = reinterpret_cast<SlabHeader*>(peer);
        const int their = side ? 0 : 1;
        hp.entries[side] = reinterpret_cast<HaloEntry*>(peer + h->mb_halo_off[their]);
        hp.count[side] = &ph->halo_count[their];
        hp.seq[side] = &ph->halo_seq[their];
      }
    k_halo_send2<<<32, 256, 0, sb>>>(S, T, h->layer_slots.as<unsigned long long>(), h->grid.as<float4>(), h->slab_lo, h->slab_hi, hp, (uint32_t)h->mb_halo_cap, seq, h->p2p_local, &S->boundary_done[0], &S->n_work[0], work_parts(h, P2G_CTAS_PER_SM));
    LAUNCH_CHECK();
  }
  if (int rc = enqueue_p2g(h, in, work_ordered(h, 0, 0))) return rc;
  stage_end(h);
  stage_begin(h, ST_HALO);
  if (has[0] || has[1]) {
    if (!concurrent) {
      // my first column goes left (the left rank holds it as halo), my halo column (== hi) goes right; a message lands in the
      // neighbour's slot for "from the right" (when I am its right neighbour) / "from the left"
      HaloPeers hp{};
      for (int side = 0; side < 2; ++side)
        if (has[side]) {
          unsigned char* peer = static_cast<unsigned char*>(h->peer_mailbox[h->rank + (side ? 1 : -1)]);
          SlabHeader* ph = reinterpret_cast<SlabHeader*>(peer);
          const int their = side ? 0 : 1;
          hp.entries[side] = reinterpret_cast<HaloEntry*>(peer + h->mb_halo_off[their]);
          hp.count[side] = &ph->halo_count[their];
          hp.seq[side] = &ph->halo_seq[their];
        }
      k_halo_send2<<<148, 256, 0, s>>>(S, T, h->layer_slots.as<unsigned long long>(), h->grid.as<float4>(), h->slab_lo, h->slab_hi, hp, (uint32_t)h->mb_halo_cap, seq, h->p2p_local, nullptr, nullptr, 1u);
      LAUNCH_CHECK();
    }
    k_halo_recv2<<<148 * 2, 256, 0, s>>>(S, T, h->layer_slots.as<unsigned long long>(), h->layer_list.as<uint32_t>(), h->grid.as<float4>(), reinterpret_cast<const HaloEntry*>(my_mb + h->mb_halo_off[0]),
                                        reinterpret_cast<const HaloEntry*>(my_mb + h->mb_halo_off[1]), my_hdr, has[0] ? 1 : 0, has[1] ? 1 : 0, seq);
    LAUNCH_CHECK();
  }
  stage_end(h);
  stage_begin(h, ST_G2P);
  CK(h->mig_list.ensure(2 * h->mb_mig_cap * 4));
  const MigrateCut cut{h->slab_lo, h->slab_hi, h->reach_lo, h->reach_hi, 4.f * (float)h->slab_lo, 4.f * (float)h->slab_hi, h->mig_list.as<uint32_t>(), h->p2p_local + 8, (uint32_t)h->mb_mig_cap};
  if (in.has_mesh)
    if (int rc = enqueue_meld(h)) return rc;
  const int src_buf = h->cur;
  BinNext bn{};
  if (bin_next) bn = BinNext{scalars_of(h, h->s_cur ^ 1), tile_table(h, h->s_cur ^ 1), bin_arrays(h, h->s_cur ^ 1)};
  DtState* D = h->dt_state.as<DtState>();
  SlabPeers peers{};
  peers.n_ranks = h->n_ranks;
  for (int side = 0; side < 2; ++side)
    if (has[side]) {
      unsigned char* peer = static_cast<unsigned char*>(h->peer_mailbox[h->rank + (side ? 1 : -1)]);
      SlabHeader* ph = reinterpret_cast<SlabHeader*>(peer);
      const int their = side ? 0 : 1;
      peers.rows[side] = reinterpret_cast<uint32_t*>(peer + h->mb_mig_off[their]);
      peers.count[side] = &ph->mig_count[their];
      peers.seq[side] = &ph->mig_seq[their];
    }
  for (int r = 0; r < h->n_ranks; ++r)
    if (r != h->rank) {
      SlabHeader* ph = reinterpret_cast<SlabHeader*>(h->peer_mailbox[r]);
      peers.err_seq[r] = &ph->err_seq[h->rank];
      peers.err_val[r] = &ph->err_val[h->rank];
    }
  // (the rows G2P writes live in the OTHER buffer until `cur` is swapped below)
  // (SVB_MIGRATE_BESIDE_G2P=0 puts the migration sender back behind G2P on the main stream, for A/B runs)
  static const bool migrate_beside = [] { const char* e = std::getenv("SVB_MIGRATE_BESIDE_G2P"); return !(e && e[0] == '0'); }();
  const bool send_beside_g2p = concurrent && !in.adaptive && migrate_beside;
  if (send_beside_g2p) {   // second stream: the migration sender, gated on the device by G2P's boundary tiles
    k_migrate_send_list<<<32, 256, 0, sb>>>(h->P(src_buf ^ 1), h->energy.as<float>(), S, cut, peers, (uint32_t)h->mb_mig_cap, seq, h->p2p_local + 10, 0, &S->boundary_done[1], &S->n_work[0], work_parts(h, G2P_CTAS_PER_SM));
    LAUNCH_CHECK();
  }
  if (!in.adaptive) {
    if (int rc = enqueue_g2p(h, in.has_mesh, /*fuse=*/true, dt, bin_next, work_ordered(h, 1, 1), src_buf, &cut)) return rc;
  } else {
    // adaptive steps: G2P with the reductions of LimitTimeStepBeforeIntegrate over all tiles, the global limits (all ranks), then
    // the advance — which notes the leavers and bins the rest ahead — as its own pass (the step is only known now)
    if (int rc = enqueue_g2p(h, in.has_mesh, /*fuse=*/false, dt, false, work_ordered(h, 1, 1), src_buf)) return rc;
    k_dt_integrate<<<1, 1, 0, s>>>(D, S, dt_peers(h));
    LAUNCH_CHECK();
    const uint32_t rows = (uint32_t)h->cap;
    if (bin_next) k_advance<true><<<blocks_for(rows, 256), 256, 0, s>>>(h->P(src_buf ^ 1), h->energy.as<float>(), S, h->K, rows, D, bn, cut);
    else k_advance<false><<<blocks_for(rows, 256), 256, 0, s>>>(h->P(src_buf ^ 1), h->energy.as<float>(), S, h->K, rows, D, bn, cut);
    LAUNCH_CHECK();
  }
  h->cur ^= 1;  // the binned buffer written by G2P is the current one from here on
  stage_end(h);
  stage_begin(h, ST_MIGRATE);
  if (!send_beside_g2p) {
    k_migrate_send_list<<<148, 256, 0, s>>>(h->Pc(), h->energy.as<float>(), S, cut, peers, (uint32_t)h->mb_mig_cap, seq, h->p2p_local + 10, 0, nullptr, nullptr, 1u);
    LAUNCH_CHECK();
  }
  CK(cudaEventRecord(h->ev_b_end, sb));
  k_migrate_recv<<<148, 256, 0, s>>>(h->Pc(), h->energy.as<float>(), S, my_hdr, reinterpret_cast<const uint32_t*>(my_mb + h->mb_mig_off[0]), reinterpret_cast<const uint32_t*>(my_mb + h->mb_mig_off[1]),
                                     has[0] ? 1 : 0, has[1] ? 1 : 0, h->rank, h->n_ranks, seq, h->n_dev, /*between_substeps=*/0, h->K, bn, bin_next ? 1 : 0, peers);
  LAUNCH_CHECK();
  if (in.adaptive) {   // the clock moves on every rank alike (after the error words of this substep have been folded in)
    k_dt_tail<<<1, 1, 0, s>>>(D, S, dt_peers(h));
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(h->h_dt, D, sizeof(DtState), cudaMemcpyDeviceToHost, s));
  }
  stage_end(h);
  ++h->lag_issued;
  h->binned_ahead = bin_next;
  h->limits_ahead = in.adaptive;
  h->have_grid = true;
  if (!in.adaptive) {
    h->time_before_last = h->time;
    h->time += (double)dt;
    ++h->substeps;
  }
  // ---- the host catches up with the front half of an earlier substep (the queued work keeps the GPU busy meanwhile)
  while (h->lag_issued - h->lag_done > (h->timing ? 0u : 2u)) {
    const int rc = process_front_p2p(h);
    if (rc) return rc;
  }
  return 0;
}

// H2D of a wire-format particle state into the current particle buffer (from_io_state, cpu/src/cpu_state.rs:26-69)
int load_particles(SvbHandle* h, const SvbParticles* p) {
  const uint32_t n = h->n;
  if (!n) return 0;
  if (!p->flags || !p->mass || !p->initial_volume || !p->mu_or_bulk_modulus || !p->lambda_or_exponent || !p->positions || !p->position_gradients || !p->velocities ||
      !p->velocity_gradients)
    return fail(h, SVB_BAD_ARGUMENT, "a required particle array is NULL");
  // stage each wire array through the spare particle buffer and transpose it into the quads
  ParticleBuf P = h->Pc();
  float* stagef = h->pbuf[h->cur ^ 1].as<float>();
  const uint32_t blocks = blocks_for(n, 256);
  // every wire array gets its own region of the spare buffer (33 n words <= 36 cap): all copies are queued back to back (the copy
  // engine never waits for a transpose), then the transposes
  struct Wire { const void* src; int field, k; float* st; };
  const Wire wires[] = {{p->flags, PFLAGS, 1, nullptr}, {p->mass, PMASS, 1, nullptr}, {p->initial_volume, PVOL, 1, nullptr}, {p->mu_or_bulk_modulus, PP0, 1, nullptr},
                        {p->lambda_or_exponent, PP1, 1, nullptr}, {p->sand_alpha, PALPHA, 1, nullptr}, {p->viscosity_dynamic, PVD, 1, nullptr}, {p->viscosity_bulk, PVB, 1, nullptr},
                        {p->collider_bits, PBITS, 1, nullptr}, {p->positions, PX, 3, nullptr}, {p->velocities, PV, 3, nullptr}, {p->velocity_gradients, PC, 9, nullptr},
                        {p->position_gradients, PF, 9, nullptr}};
  Wire queued[sizeof(wires) / sizeof(wires[0])];
  int n_queued = 0;
  size_t stage_off = 0;
  for (const Wire& w : wires) {
    if (!w.src) continue;
    Wire q = w;
    q.st = stagef + stage_off;
    stage_off += (size_t)h->cap * w.k;
    CK(cudaMemcpyAsync(q.st, w.src, (size_t)n * w.k * 4, cudaMemcpyHostToDevice, h->stream));
    queued[n_queued++] = q;
  }
  for (int j = 0; j < n_queued; ++j) {
    const Wire& q = queued[j];
    if (q.k == 1) k_wire_to_soa<1><<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const uint32_t*>(q.st), P, q.field, n);
    else if (q.k == 3) k_wire_to_soa<3><<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const uint32_t*>(q.st), P, q.field, n);
    else k_wire_to_soa<9><<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const uint32_t*>(q.st), P, q.field, n);
    LAUNCH_CHECK();
  }
  if (p->elastic_energies) CK(cudaMemcpyAsync(h->energy.p, p->elastic_energies, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream));
  k_iota_orig<<<blocks, 256, 0, h->stream>>>(P, n, 0);
  LAUNCH_CHECK();
  // (host work behind the queued copies: it overlaps the H2D transfers)
  if (p->initial_positions) h->initial_positions.assign(p->initial_positions, p->initial_positions + (size_t)n * 3);
  else h->initial_positions.assign((size_t)n * 3, 0.f);
  return 0;
}

int read_status(SvbHandle* h) {
  StepScalars* S = cur_scalars(h);
  CK(cudaMemcpyAsync(h->h_scalars, S, sizeof(StepScalars), cudaMemcpyDeviceToHost, h->stream));
  CK(sync_streams(h));
  h->status |= (h->h_scalars->status | h->h_scalars->sticky | h->h_scalars->sticky_new | h->h_scalars->accum) & 0xffffu;
  if ((h->h_scalars->sticky_new & 0xffffu) && !h->adaptive.has_override) rollback_failed_substep(h);   // fixed dt: the failing substep was the last one queued
  return 0;
}

}  // namespace

extern "C" {

int32_t svb_available_devices(char* out, size_t cap) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) return SVB_CUDA_ERROR;
  std::string s;
  for (int d = 0; d < count; ++d) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) != cudaSuccess) return SVB_CUDA_ERROR;
    s += std::to_string(d) + ": " + p.name + "\n";
  }
  if (out && cap) {
    const size_t m = std::min(cap - 1, s.size());
    std::memcpy(out, s.data(), m);
    out[m] = 0;
  }
  return count;
}

int32_t svb_create(const SvbConsts* consts, const SvbParticles* p, double time, int32_t device, SvbHandle** out) {
  if (!consts || !p || !out) return SVB_BAD_ARGUMENT;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SVB_CUDA_ERROR;  // no CPU fallback
  if (device < 0 || device >= count) return SVB_BAD_ARGUMENT;
  if (p->n > 0xfffffff0ull) return SVB_BAD_ARGUMENT;
  SvbHandle* h = new SvbHandle();
  *out = h;  // returned even on failure so the caller can read svb_last_error, then svb_destroy
  h->device = device;
  h->consts = *consts;
  h->K.h = consts->grid_node_size / consts->simulation_scale;                 // header.rs:36-38
  h->K.leaf_size = consts->leaf_size;                                         // unscaled
  h->K.accept_distance = h->K.h * 2.f;                                        // header.rs:60-62
  h->K.forget_distance = h->K.h * 2.2f;                                       // header.rs:64-66
  for (int k = 0; k < 3; ++k) {
    h->K.domain_min[k] = consts->domain_min[k] / consts->simulation_scale;    // header.rs:40-58
    h->K.domain_max[k] = consts->domain_max[k] / consts->simulation_scale;
  }
  h->time = time;
  CK(cudaSetDevice(device));
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    int lo_prio = 0, hi_prio = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    CK(cudaStreamCreateWithPriority(&h->stream_b, cudaStreamNonBlocking, hi_prio));   // its few blocks should get an SM slot as soon as they are launched
    CK(cudaEventCreateWithFlags(&h->ev_b_start, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_b_end, cudaEventDisableTiming));
    CK(cudaEventRecord(h->ev_b_end, h->stream_b));
  }
  for (auto& e : h->ev) CK(cudaEventCreate(&e));
  for (auto& e : h->ev_adv) CK(cudaEventCreate(&e));
  for (auto& L : h->lag) {
    CK(cudaMallocHost(&L.host, sizeof(StepScalars)));
    CK(cudaEventCreateWithFlags(&L.ev, cudaEventDisableTiming));
  }
  h->h_scalars = h->lag[0].host;
  h->ev_front = h->lag[0].ev;
  CK(cudaMallocHost(&h->h_dt, sizeof(DtState)));
  CK(h->dt_state.ensure(sizeof(DtState)));
  CK(cudaMemsetAsync(h->dt_state.p, 0, sizeof(DtState), h->stream));
  if (int rc = preload_kernels(h)) return rc;
  CK(cudaFuncSetAttribute(k_p2g<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2G_SMEM));
  CK(cudaFuncSetAttribute(k_p2g<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2G_SMEM));
  const uint32_t n = (uint32_t)p->n;
  h->n = n;
  h->cap = ((size_t)std::max<uint32_t>(n, 1) + 63) & ~(size_t)63;
  for (int b = 0; b < 2; ++b) {
    CK(h->pbuf[b].ensure(h->cap * NWORDS * 4));
  }
  CK(h->energy.ensure(h->cap * 4));
  CK(h->scalars.ensure(2 * sizeof(StepScalars)));
  CK(cudaMemsetAsync(h->scalars.p, 0, 2 * sizeof(StepScalars), h->stream));
  CK(h->pcell.ensure(h->cap * 4));
  CK(h->prank.ensure(h->cap * 4));
  CK(h->src_of.ensure(h->cap * 4));
  CK(h->scratch.ensure(4096));
  CK(h->layer_slots.ensure(LAYER_SLOTS * 8));
  CK(h->layer_list.ensure(LAYER_SLOTS * 4));
  CK(cudaMemsetAsync(h->layer_slots.p, 0, LAYER_SLOTS * 8, h->stream));
  for (int b = 0; b < 2; ++b) CK(cudaMemsetAsync(h->pbuf[b].p, 0, h->cap * NWORDS * 4, h->stream));
  CK(cudaMemsetAsync(h->energy.p, 0, h->cap * 4, h->stream));
  if (int rc = ensure_tile_capacity(h, (size_t)n / 96 + 2048)) return rc;

  if (int rc = load_particles(h, p)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  // an empty collider set is a valid input (the reference builds an empty Topology)
  h->topo = svbh::HostTopology();
  h->have_topology = true;
  return 0;
}

void svb_destroy(SvbHandle* h) {
  if (!h) return;
  if (h->multi) return multi_destroy(h);
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (auto& f : h->fs)
    for (DevBuf* b : {&f.table_slots, &f.tile_key, &f.tile_slot, &f.tile_touch, &f.cell_count, &f.slot_first, &f.tile_start, &f.nbr}) b->release();
  if (h->h_dt) cudaFreeHost(h->h_dt);
  DevBuf* all[] = {&h->pbuf[0], &h->pbuf[1], &h->energy, &h->pcell, &h->prank, &h->src_of, &h->dt_state,
                   &h->grid, &h->melded, &h->mig_list, &h->work_list, &h->node_mask, &h->node_offset, &h->scratch, &h->scalars, &h->layer_slots, &h->layer_list,
                   &h->d_tri, &h->d_opp, &h->d_tri_collider, &h->d_fan_offsets, &h->d_fan_tris, &h->d_va, &h->d_vb, &h->d_vvel, &h->d_fric_a, &h->d_fric_b, &h->d_damp_a, &h->d_damp_b,
                   &h->d_vpos, &h->d_vnormal, &h->d_tnormal, &h->d_tbox, &h->d_tfric, &h->d_tdamp, &h->d_node_min, &h->d_node_max, &h->d_node_first, &h->d_node_count, &h->d_children,
                   &h->d_tri_indices, &h->d_flags_a, &h->d_flags_b, &h->d_goal_a, &h->d_goal_b, &h->snap_p, &h->snap_e,
                   &h->comm_counts, &h->mailbox, &h->halo_send[0], &h->halo_send[1], &h->halo_recv[0], &h->halo_recv[1], &h->mig_send[0], &h->mig_send[1], &h->mig_recv[0], &h->mig_recv[1]};
  for (DevBuf* b : all) b->release();
  for (auto& L : h->lag) {
    if (L.host) cudaFreeHost(L.host);
    if (L.ev) cudaEventDestroy(L.ev);
  }
  for (void* pm : h->peer_mailbox)
    if (pm && h->ipc_mapped) cudaIpcCloseMemHandle(pm);
  if (h->comm) ncclCommDestroy(h->comm);
  if (h->h_counts) cudaFreeHost(h->h_counts);
  for (auto& e : h->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : h->ev_adv)
    if (e) cudaEventDestroy(e);
  if (h->stream_b) { cudaStreamSynchronize(h->stream_b); cudaStreamDestroy(h->stream_b); }
  if (h->ev_b_start) cudaEventDestroy(h->ev_b_start);
  if (h->ev_b_end) cudaEventDestroy(h->ev_b_end);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int32_t svb_upload(SvbHandle* h, const SvbParticles* p, double time) {
  if (!h || !p) return SVB_BAD_ARGUMENT;
  if (h->multi) return multi_upload(h, p, time);
  if (p->n > 0xfffffff0ull) return SVB_BAD_ARGUMENT;
  if (int rc = set_device(h)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  const uint32_t n = (uint32_t)p->n;
  const size_t need = h->slabs ? (size_t)n * 3 / 2 + 65536 : (size_t)std::max<uint32_t>(n, 1);
  if (int rc = resize_particles(h, need)) return rc;
  h->n = n;
  for (int b = 0; b < 2; ++b) CK(cudaMemsetAsync(h->pbuf[b].p, 0, h->cap * NWORDS * 4, h->stream));
  CK(cudaMemsetAsync(h->energy.p, 0, h->cap * 4, h->stream));
  if (int rc = load_particles(h, p)) return rc;
  // a new state starts a new run: clock, step history, error words, per-substep tables
  h->time = time;
  h->time_before_last = time;
  h->failed_rolled_back = false;
  h->substeps = 0;
  h->status = 0;
  h->adaptive = svbh::AdaptiveTimeStep();
  // (the scalars keep the tile counts of each front set: the next reset undoes exactly those)
  h->binned_ahead = false;
  h->limits_ahead = false;
  h->n_ptiles = h->n_live = h->n_tiles = 0;
  h->have_grid = false;
  h->masks_valid = false;
  h->have_snapshot = false;
  if (h->slabs) {
    if (n && h->orig_offset) {
      k_add_orig<<<blocks_for(n, 256), 256, 0, h->stream>>>(h->Pc(), n, (uint32_t)h->orig_offset);
      LAUNCH_CHECK();
    }
    if (h->p2p) CK(cudaMemcpyAsync(h->n_dev, &h->n, 4, cudaMemcpyHostToDevice, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int32_t svb_set_topology(SvbHandle* h, uint32_t n_colliders, const uint32_t* num_vertices, const uint32_t* num_triangles, const uint32_t* triangles) {
  if (!h) return SVB_BAD_ARGUMENT;
  if (h->multi) return multi_set_topology(h, n_colliders, num_vertices, num_triangles, triangles);
  if (n_colliders > 16) return fail(h, SVB_TOO_MANY_COLLIDERS, "too many colliders: %u (at most 16)", n_colliders);
  if (int rc = set_device(h)) return rc;
  const std::string err = h->topo.build(n_colliders, num_vertices, num_triangles, triangles);
  if (!err.empty()) {
    h->topo = svbh::HostTopology();
    return fail(h, SVB_BAD_MESH, "Something is wrong with the mesh inputs: %s", err.c_str());
  }
  const auto& T = h->topo;
  int rc = 0;
  if ((rc = upload_array(h, h->d_tri, T.tri.data(), T.tri.size() * 4)) || (rc = upload_array(h, h->d_opp, T.opp.data(), T.opp.size() * 4)) ||
      (rc = upload_array(h, h->d_tri_collider, T.tri_collider.data(), T.tri_collider.size() * 4)) ||
      (rc = upload_array(h, h->d_fan_offsets, T.fan_offsets.data(), T.fan_offsets.size() * 4)) || (rc = upload_array(h, h->d_fan_tris, T.fan_tris.data(), T.fan_tris.size() * 4)))
    return rc;
  CK(h->d_vpos.ensure((size_t)T.n_vertices * 12 + 4));
  CK(h->d_vnormal.ensure((size_t)T.n_vertices * 12 + 4));
  CK(h->d_tnormal.ensure((size_t)T.n_triangles * 12 + 4));
  CK(h->d_tbox.ensure((size_t)T.n_triangles * 32 + 16));
  CK(h->d_tfric.ensure((size_t)T.n_triangles * 4 + 4));
  CK(h->d_tdamp.ensure((size_t)T.n_triangles * 4 + 4));
  CK(cudaStreamSynchronize(h->stream));
  h->have_topology = true;
  h->have_keyframes = false;
  return 0;
}

int32_t svb_set_keyframes(SvbHandle* h, uint64_t frame, const SvbKeyframe* a, const SvbKeyframe* b) {
  if (!h || !a) return SVB_BAD_ARGUMENT;
  if (h->multi) return multi_set_keyframes(h, frame, a, b);
  if (int rc = set_device(h)) return rc;
  const auto& T = h->topo;
  const uint32_t nv = T.n_vertices, nt = T.n_triangles, n = h->n_global ? h->n_global : h->n;  // goal arrays are in (global) original order
  if (nv && (!a->vertex_positions || (b && !b->vertex_positions))) return fail(h, SVB_BAD_ARGUMENT, "keyframe lacks vertex_positions");
  h->frame = frame;
  h->has_b = b != nullptr;
  for (int k = 0; k < 3; ++k) {
    h->gravity_a[k] = a->gravity[k];
    h->gravity_b[k] = b ? b->gravity[k] : a->gravity[k];
  }
  int rc = 0;
  // goals (external_force.rs:38-41): only uploaded when some particle has HAS_GOAL in keyframe a
  h->has_goals = false;
  if (a->particle_flags && a->particle_goal_positions && n) {
    bool any = false;
    for (uint32_t i = 0; i < n && !any; ++i) any = (a->particle_flags[i] & F_HAS_GOAL) != 0;
    if (any) {
      if ((rc = upload_array(h, h->d_flags_a, a->particle_flags, (size_t)n * 4)) || (rc = upload_array(h, h->d_goal_a, a->particle_goal_positions, (size_t)n * 12))) return rc;
      if (b) {
        if (!b->particle_flags || !b->particle_goal_positions) return fail(h, SVB_BAD_ARGUMENT, "keyframe b lacks goal arrays");
        if ((rc = upload_array(h, h->d_flags_b, b->particle_flags, (size_t)n * 4)) || (rc = upload_array(h, h->d_goal_b, b->particle_goal_positions, (size_t)n * 12))) return rc;
      }
      h->has_goals = true;
    }
  }
  MeshDev M{};
  M.n_vertices = nv; M.n_triangles = nt; M.n_colliders = T.n_colliders;
  if (nt) {
    std::vector<float> zeros(nt, 0.f);
    const float* fa = a->triangle_frictions ? a->triangle_frictions : zeros.data();
    const float* da = a->triangle_dampings ? a->triangle_dampings : zeros.data();
    const float* fb = b ? (b->triangle_frictions ? b->triangle_frictions : zeros.data()) : fa;
    const float* db = b ? (b->triangle_dampings ? b->triangle_dampings : zeros.data()) : da;
    const float* va = a->vertex_positions;
    const float* vb = b ? b->vertex_positions : a->vertex_positions;
    // linear_vertex_velocities (xpu/src/frame_input.rs:334-348)
    std::vector<float> vvel((size_t)nv * 3, 0.f);
    if (b)
      for (size_t i = 0; i < (size_t)nv * 3; ++i) vvel[i] = (vb[i] - va[i]) * (float)h->consts.frames_per_second;
    if ((rc = upload_array(h, h->d_va, va, (size_t)nv * 12)) || (rc = upload_array(h, h->d_vb, vb, (size_t)nv * 12)) || (rc = upload_array(h, h->d_vvel, vvel.data(), (size_t)nv * 12)) ||
        (rc = upload_array(h, h->d_fric_a, fa, (size_t)nt * 4)) || (rc = upload_array(h, h->d_fric_b, fb, (size_t)nt * 4)) || (rc = upload_array(h, h->d_damp_a, da, (size_t)nt * 4)) ||
        (rc = upload_array(h, h->d_damp_b, db, (size_t)nt * 4)))
      return rc;
    // update_bvh (xpu/src/frame_input.rs:350-390)
    h->bvh = svbh::BvhBuilder::build(T, va, b ? vb : nullptr, h->K.forget_distance, h->consts.leaf_size, h->consts.leaf_threshold);
    const auto& B = h->bvh;
    if ((rc = upload_array(h, h->d_node_min, B.node_min.data(), B.node_min.size() * 4)) || (rc = upload_array(h, h->d_node_max, B.node_max.data(), B.node_max.size() * 4)) ||
        (rc = upload_array(h, h->d_node_first, B.node_first.data(), B.node_first.size() * 4)) || (rc = upload_array(h, h->d_node_count, B.node_count.data(), B.node_count.size() * 4)) ||
        (rc = upload_array(h, h->d_children, B.children.data(), B.children.size() * 4)) || (rc = upload_array(h, h->d_tri_indices, B.tri_indices.data(), B.tri_indices.size() * 4)))
      return rc;
    CK(cudaStreamSynchronize(h->stream));  // the host vectors above go out of scope
    M.tri = h->d_tri.as<uint32_t>(); M.opp = h->d_opp.as<uint32_t>(); M.tri_collider = h->d_tri_collider.as<uint32_t>();
    M.fan_offsets = h->d_fan_offsets.as<uint32_t>(); M.fan_tris = h->d_fan_tris.as<uint32_t>();
    M.va = h->d_va.as<float>(); M.vb = h->d_vb.as<float>(); M.vvel = h->d_vvel.as<float>();
    M.fric_a = h->d_fric_a.as<float>(); M.fric_b = h->d_fric_b.as<float>(); M.damp_a = h->d_damp_a.as<float>(); M.damp_b = h->d_damp_b.as<float>();
    M.vpos = h->d_vpos.as<float>(); M.vnormal = h->d_vnormal.as<float>(); M.tnormal = h->d_tnormal.as<float>(); M.tbox = h->d_tbox.as<float>(); M.tfric = h->d_tfric.as<float>(); M.tdamp = h->d_tdamp.as<float>();
    M.bvh_level = B.level; M.bvh_nodes = (int32_t)B.node_count.size();
    M.node_min = h->d_node_min.as<int32_t>(); M.node_max = h->d_node_max.as<int32_t>(); M.node_first = h->d_node_first.as<int32_t>(); M.node_count = h->d_node_count.as<int32_t>();
    M.children = h->d_children.as<int32_t>(); M.tri_indices = h->d_tri_indices.as<uint32_t>();
  }
  CK(cudaStreamSynchronize(h->stream));
  h->M = M;
  h->have_keyframes = true;
  return 0;
}

int32_t svb_advance(SvbHandle* h, double target_time, float max_time_step, int32_t adaptive_time_steps, const volatile int32_t* cancel, void (*progress)(void*, size_t), void* user) {
  if (!h) return SVB_BAD_ARGUMENT;
  if (h->multi) return multi_advance(h, target_time, max_time_step, adaptive_time_steps, cancel, progress, user);
  if (!h->have_keyframes) return fail(h, SVB_INPUT_MISSING, "At this point, interpolated input should be ready (svb_set_keyframes not called)");
  if (int rc = set_device(h)) return rc;
  const bool adaptive = adaptive_time_steps != 0;
  cudaStream_t s = h->stream;
  h->adaptive.max_time_step = max_time_step;
  h->status = 0;
  h->stop_bits = 0;
  {  // a new advance starts without status bits; the tile bookkeeping of the last substep stays (the next reset undoes it)
    for (int k = 0; k < 2; ++k) {
      StepScalars* S = scalars_of(h, k);
      CK(cudaMemsetAsync(&S->sticky, 0, 12, s));   // sticky, sticky_new, accum  (the words are adjacent, see svb_device.cuh)
    }
    CK(cudaMemsetAsync(&cur_scalars(h)->status, 0, 4, s));   // (the other half may hold what the binning-ahead raised: that stays)
  }
  h->failed_rolled_back = false;
  if (h->timing) std::memset(h->stage_ms, 0, sizeof h->stage_ms);
  const double spf = 1.0 / (double)h->consts.frames_per_second;
  CK(cudaEventRecord(h->ev_adv[0], s));
  // room for the per-substep inflow (a trickle: the CFL limit keeps travel below one cell per substep); the loop itself never
  // resizes, k_migrate_recv reports a full buffer.  Deliberately NOT sized by the mailboxes (those also serve a rebalance, which
  // makes its own room): slab ranks launch their per-row kernels for the capacity, idle blocks are not free.
  const size_t inflow = (size_t)h->n / 16 + 65536;
  if (h->p2p && (size_t)h->n + 2 * inflow > h->cap) {
    if (int rc = resize_particles(h, ((size_t)h->n + 2 * inflow) * 5 / 4)) return rc;
  }
  h->device_clock = adaptive && (h->slabs ? h->p2p : h->n > 0);
  if (h->device_clock) {
    // ---- the clock and AdaptiveTimeStepState move to the device for the duration of the call
    DtState d{};
    d.time = h->time; d.target = target_time; d.fps = (double)h->consts.frames_per_second; d.frame = h->frame;
    d.max_dt = max_time_step; d.h = h->K.h;
    const auto& a = h->adaptive;
    d.has = (a.has_velocity ? 1u : 0u) | (a.has_deformation ? 2u : 0u) | (a.has_isolated ? 4u : 0u) | (a.has_sound ? 8u : 0u);
    d.by_velocity = a.by_velocity; d.by_deformation = a.by_deformation; d.by_isolated = a.by_isolated; d.by_sound = a.by_sound;
    d.prior_len = (uint32_t)std::min<size_t>(a.prior.size(), 11);
    for (uint32_t q = 0; q < d.prior_len; ++q) d.prior[q] = a.prior[q];
    for (int k = 0; k < 3; ++k) { d.ga[k] = h->gravity_a[k]; d.gb[k] = h->has_b ? h->gravity_b[k] : h->gravity_a[k]; }
    d.next_min_sound_key = INT32_MAX; d.next_min_isolated_key = INT32_MAX; d.next_live = 0;
    DtState* D = h->dt_state.as<DtState>();
    if (h->limits_ahead) {   // the limits of the current state were reduced by the last k_advance: keep them
      DtState old{};
      CK(cudaMemcpyAsync(&old, D, sizeof old, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      d.next_min_sound_key = old.next_min_sound_key; d.next_min_isolated_key = old.next_min_isolated_key; d.next_live = old.next_live;
    }
    *h->h_dt = d;
    CK(cudaMemcpyAsync(D, h->h_dt, sizeof d, cudaMemcpyHostToDevice, s));
    if (!h->limits_ahead) {
      const uint32_t rows = h->p2p ? (uint32_t)h->cap : h->n;
      k_limit_force<<<std::max<uint32_t>(blocks_for(rows, 256), 1), 256, 0, s>>>(h->Pc(), D, h->K.h, rows, h->p2p ? h->n_dev : nullptr);
      LAUNCH_CHECK();
    }
    k_dt_open<<<1, 1, 0, s>>>(D, cur_scalars(h), dt_peers(h));
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(h->h_dt, D, sizeof d, cudaMemcpyDeviceToHost, s));
    h->limits_ahead = false;
    int rc = 0;
    for (;;) {
      if (cancel && *cancel) { rc = fail(h, SVB_CANCELED, "The computation was canceled"); break; }
      rc = substep(h, true);
      if (rc != 0) break;
      // (h_dt is the device clock as of the previous substep's tail: the wait for this substep's front half has passed it)
      if (progress) progress(user, (size_t)(std::fmod(h->h_dt->time, spf) * 1000.0));
    }
    if (rc < 0) return rc;
    while (h->p2p && h->lag_done < h->lag_issued) {   // (a slab rank may not have looked at its last fronts yet)
      const int rc2 = process_front_p2p(h);
      if (rc2 < 0) return rc2;
      if (rc2 == 2) break;
    }
    // rc == 2: a stop word ended the run — the target time, a failed particle, a wrong frame, a zero time step
    CK(cudaMemcpyAsync(h->h_dt, D, sizeof d, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const DtState& e = *h->h_dt;
    h->time = e.time;
    h->substeps += e.substeps;
    auto& w = h->adaptive;
    w.has_velocity = e.has & 1u; w.has_deformation = e.has & 2u; w.has_isolated = e.has & 4u; w.has_sound = e.has & 8u;
    w.by_velocity = e.by_velocity; w.by_deformation = e.by_deformation; w.by_isolated = e.by_isolated; w.by_sound = e.by_sound;
    w.prior.assign(e.prior, e.prior + std::min<uint32_t>(e.prior_len, 11));
    w.allowed_override = e.allowed; w.has_override = true;
    h->limits_ahead = (h->stop_bits & ST_STOP_DONE) != 0 && !(h->stop_bits & 0xffffu);   // the pending LimitTimeStepBeforeForce belongs to the next call
    if (h->stop_bits & ST_STOP_ZERO_DT) return fail(h, SVB_ZERO_TIME_STEP, "The time step ended up being 0");
    if (h->stop_bits & ST_STOP_FRAME) return fail(h, SVB_FRAME_INPUT, "Wrong frame loaded: %llu (need %llu)", (unsigned long long)h->frame, (unsigned long long)std::floor(e.time * e.fps));
  } else {
    if (adaptive && h->slabs) return fail(h, SVB_BAD_ARGUMENT, "adaptive time steps on slab ranks need the peer-memory exchange path (SVB_SLAB_NCCL forces the NCCL fallback)");
    h->adaptive.has_override = false;
    if (adaptive && h->n == 0) {   // nothing limits the step: the reference walks to the target in steps of max_time_step
      h->adaptive = svbh::AdaptiveTimeStep();
      h->adaptive.max_time_step = max_time_step;
    }
    while (h->time < target_time) {
      if (cancel && *cancel) return fail(h, SVB_CANCELED, "The computation was canceled");
      const int rc = substep(h, false);
      if (rc < 0) return rc;
      if (rc == 2 || (h->status & SVB_PARTICLE_CLOSE_TO_INVERTED)) break;
      if (progress) progress(user, (size_t)(std::fmod(h->time, spf) * 1000.0));
    }
    while (h->p2p && h->lag_done < h->lag_issued) {   // the fronts of the last substeps have not been looked at yet
      const int rc = process_front_p2p(h);
      if (rc < 0) return rc;
      if (rc == 2) break;
    }
  }
  CK(cudaEventRecord(h->ev_adv[1], s));
  if (int rc = read_status(h)) return rc;
  if (h->p2p) {  // the row count lived on the device during the loop
    CK(cudaMemcpy(&h->n, h->n_dev, 4, cudaMemcpyDeviceToHost));
    if (h->h_scalars->status & (ST_COMM_TIMEOUT | ST_COMM_OVERFLOW)) {
      const StepScalars& r = *h->h_scalars;
      return fail(h, SVB_COMM_ERROR, "%s (rank %d, message %u; boundary tiles %u, interior %u, ticks P2G %u G2P %u, status 0x%x)",
                  r.status & ST_COMM_TIMEOUT ? "a neighbour slab's message did not arrive" : "a slab mailbox or the particle buffer ran out of room", h->rank, h->slab_seq, r.n_work[0],
                  r.n_work[1], r.boundary_done[0], r.boundary_done[1], r.status);
    }
  }
  CK(cudaEventElapsedTime(&h->last_advance_ms, h->ev_adv[0], h->ev_adv[1]));
  if (h->status) {
    if (h->status & SVB_PARTICLE_CLOSE_TO_INVERTED) fail(h, 0, "Failed to compute the elastic energy of a particle (EnergyError::PositionGradientNonPositive)");
    else fail(h, 0, "device status word 0x%x", h->status);
    return (int32_t)h->status;
  }
  return 0;
}

int32_t svb_download(SvbHandle* h, SvbParticles* out) {
  if (!h || !out) return SVB_BAD_ARGUMENT;
  if (h->multi) return multi_download(h, out);
  // slab ranks hold rows of the GLOBAL particle order (and F_GONE rows): the original-order scatter below would leave its buffers
  if (h->slabs) return fail(h, SVB_BAD_ARGUMENT, "svb_download needs the whole particle set on one device: use svb_download_resident on a slab rank");
  if (int rc = set_device(h)) return rc;
  const uint32_t n = h->n;
  out->n = n;
  if (!n) return 0;
  ParticleBuf P = h->Pc();
  float* stagef = h->pbuf[h->cur ^ 1].as<float>();  // the spare buffer is free between substeps
  const uint32_t blocks = blocks_for(n, 256);
  // one region of the spare buffer per field (33 n words in total): all gathers into original order are queued first, then the
  // copies back to back (the copy engine never waits for a gather)
  struct Wire { void* dst; int word, k; float* st; };
  const Wire wires[] = {{out->flags, PFLAGS, 1, nullptr}, {out->mass, PMASS, 1, nullptr}, {out->initial_volume, PVOL, 1, nullptr}, {out->mu_or_bulk_modulus, PP0, 1, nullptr},
                        {out->lambda_or_exponent, PP1, 1, nullptr}, {out->sand_alpha, PALPHA, 1, nullptr}, {out->viscosity_dynamic, PVD, 1, nullptr}, {out->viscosity_bulk, PVB, 1, nullptr},
                        {out->collider_bits, PBITS, 1, nullptr}, {out->positions, PX, 3, nullptr}, {out->velocities, PV, 3, nullptr}, {out->velocity_gradients, PC, 9, nullptr},
                        {out->position_gradients, PF, 9, nullptr}};
  Wire queued[sizeof(wires) / sizeof(wires[0])];
  int n_queued = 0;
  size_t stage_off = 0;
  for (const Wire& w : wires) {
    if (!w.dst) continue;
    Wire q = w;
    q.st = stagef + stage_off;
    stage_off += (size_t)h->cap * w.k;
    if (q.k == 1) k_soa_to_wire<1><<<blocks, 256, 0, h->stream>>>(P, q.word, reinterpret_cast<uint32_t*>(q.st), n);
    else if (q.k == 3) k_soa_to_wire<3><<<blocks, 256, 0, h->stream>>>(P, q.word, reinterpret_cast<uint32_t*>(q.st), n);
    else k_soa_to_wire<9><<<blocks, 256, 0, h->stream>>>(P, q.word, reinterpret_cast<uint32_t*>(q.st), n);
    LAUNCH_CHECK();
    queued[n_queued++] = q;
  }
  // energies live in their own array: park them in the node-mask scratch
  if (out->elastic_energies) {
    CK(h->scratch.ensure((size_t)h->cap * 4));
    k_array_to_wire<<<blocks, 256, 0, h->stream>>>(h->energy.as<float>(), P, h->scratch.as<float>(), n);
    LAUNCH_CHECK();
  }
  for (int j = 0; j < n_queued; ++j) CK(cudaMemcpyAsync(queued[j].dst, queued[j].st, (size_t)n * queued[j].k * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out->elastic_energies) CK(cudaMemcpyAsync(out->elastic_energies, h->scratch.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out->initial_positions) std::memcpy(out->initial_positions, h->initial_positions.data(), (size_t)n * 12);   // while the D2H copies run
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int build_node_masks(SvbHandle* h, uint32_t* total) {
  *total = 0;
  if (!h->have_grid || h->n_tiles == 0) return 0;
  const uint32_t na = h->n_tiles;
  if (!h->masks_valid) {
    // store_grid was off during the last substep: fall back to "node holds mass or momentum"
    k_mask_from_values<<<na, 64, 0, h->stream>>>(h->grid.as<float4>(), h->node_mask.as<unsigned long long>());
    LAUNCH_CHECK();
  }
  k_popc_masks<<<blocks_for(na, 256), 256, 0, h->stream>>>(h->node_mask.as<unsigned long long>(), h->node_offset.as<uint32_t>(), na);
  LAUNCH_CHECK();
  k_scan_tiles<<<1, 1024, 0, h->stream>>>(h->node_offset.as<uint32_t>(), nullptr, na, h->scratch.as<uint32_t>());
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(total, h->scratch.p, 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

// The grid is the one of the last substep's P2G (positions before the last advance), like the
// reference's `grid_nodes` after produce_next_state.  Emitted nodes = nodes with >= 1 contributor
// (exact when the "store_grid" option was on during the substep; the reference additionally keeps
// zero-mass nodes touched one substep earlier, SURVEY.md §8g.9 — those are not emitted).
int64_t svb_grid_count(SvbHandle* h) {
  if (!h) return SVB_BAD_ARGUMENT;
  if (h->multi) return fail(h, SVB_BAD_ARGUMENT, "single-device introspection entry point called on a multi-device handle");
  if (int rc = set_device(h)) return rc;
  uint32_t total = 0;
  if (int rc = build_node_masks(h, &total)) return rc;
  return (int64_t)total;
}

int32_t svb_download_grid(SvbHandle* h, SvbGrid* out) {
  if (!h || !out) return SVB_BAD_ARGUMENT;
  if (h->multi) return fail(h, SVB_BAD_ARGUMENT, "single-device introspection entry point called on a multi-device handle");
  if (int rc = set_device(h)) return rc;
  uint32_t total = 0;
  if (int rc = build_node_masks(h, &total)) return rc;
  if (out->n < total) return fail(h, SVB_BAD_ARGUMENT, "svb_download_grid: capacity %llu < %u nodes", (unsigned long long)out->n, total);
  out->n = total;
  if (!total) return 0;
  DevBuf ids, bits, masses, vels;
  CK(ids.ensure((size_t)total * 12));
  CK(bits.ensure((size_t)total * 4));
  CK(masses.ensure((size_t)total * 4));
  CK(vels.ensure((size_t)total * 12));
  k_emit_grid<<<h->n_tiles, 64, 0, h->stream>>>(h->grid.as<float4>(), tile_table(h), meld_info(h), cur_scalars(h), h->node_mask.as<unsigned long long>(),
                                              h->node_offset.as<uint32_t>(), ids.as<int32_t>(), bits.as<uint32_t>(), masses.as<float>(), vels.as<float>());
  LAUNCH_CHECK();
  if (out->node_ids) CK(cudaMemcpyAsync(out->node_ids, ids.p, (size_t)total * 12, cudaMemcpyDeviceToHost, h->stream));
  if (out->collider_bits) CK(cudaMemcpyAsync(out->collider_bits, bits.p, (size_t)total * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out->masses) CK(cudaMemcpyAsync(out->masses, masses.p, (size_t)total * 4, cudaMemcpyDeviceToHost, h->stream));
  if (out->velocities) CK(cudaMemcpyAsync(out->velocities, vels.p, (size_t)total * 12, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (out->contributor_counts)
    for (uint32_t i = 0; i < total; ++i) out->contributor_counts[i] = 1;
  ids.release(); bits.release(); masses.release(); vels.release();
  return 0;
}

double svb_time(const SvbHandle* h) { return h ? h->time : 0.0; }
uint64_t svb_substeps(const SvbHandle* h) { return h ? h->substeps : 0; }
float svb_allowed_time_step(const SvbHandle* h) { return h ? h->adaptive.allowed() : 0.f; }
uint32_t svb_status(const SvbHandle* h) { return h ? h->status : 0; }
const char* svb_last_error(const SvbHandle* h) { return h ? h->last_error.c_str() : "null handle"; }
uint64_t svb_kernel_launches(const SvbHandle* h) { return h ? h->launches : 0; }
float svb_last_advance_ms(const SvbHandle* h) { return h ? h->last_advance_ms : 0.f; }
uint64_t svb_particle_count(const SvbHandle* h) { return h ? h->n : 0; }

int32_t svb_binning(SvbHandle* h, uint32_t* sort_map, int32_t* cells) {
  if (!h) return SVB_BAD_ARGUMENT;
  if (h->multi) return fail(h, SVB_BAD_ARGUMENT, "single-device introspection entry point called on a multi-device handle");
  if (int rc = set_device(h)) return rc;
  const uint32_t n = h->n;
  if (!n) return 0;
  ParticleBuf P = h->Pc();
  if (sort_map) {
    uint32_t* d = h->pbuf[h->cur ^ 1].as<uint32_t>() + (size_t)3 * h->cap;   // behind the cells
    k_soa_to_wire_plain<<<blocks_for(n, 256), 256, 0, h->stream>>>(P, PORIG, d, n);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(sort_map, d, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  if (cells) {
    int32_t* d = h->pbuf[h->cur ^ 1].as<int32_t>();
    k_cells<<<blocks_for(n, 256), 256, 0, h->stream>>>(P, h->K.h, n, d);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(cells, d, (size_t)n * 12, cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int64_t svb_active_block_count(SvbHandle* h) { return h ? (h->have_grid ? (int64_t)h->n_tiles : 0) : SVB_BAD_ARGUMENT; }

int32_t svb_active_blocks(SvbHandle* h, int32_t* block_ids, uint32_t* collider_bits) {
  if (!h) return SVB_BAD_ARGUMENT;
  if (h->multi) return fail(h, SVB_BAD_ARGUMENT, "single-device introspection entry point called on a multi-device handle");
  if (int rc = set_device(h)) return rc;
  const uint32_t na = h->have_grid ? h->n_tiles : 0;
  if (!na) return 0;
  DevBuf ids, bits;
  CK(ids.ensure((size_t)na * 12));
  CK(bits.ensure((size_t)na * 4));
  k_decode_active<<<blocks_for(na, 256), 256, 0, h->stream>>>(tile_table(h), h->layer_slots.as<unsigned long long>(), na, ids.as<int32_t>(), bits.as<uint32_t>());
  LAUNCH_CHECK();
  if (block_ids) CK(cudaMemcpyAsync(block_ids, ids.p, (size_t)na * 12, cudaMemcpyDeviceToHost, h->stream));
  if (collider_bits) CK(cudaMemcpyAsync(collider_bits, bits.p, (size_t)na * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  ids.release(); bits.release();
  return 0;
}

int32_t svb_stage_times(SvbHandle* h, const char** names, float* ms, int32_t cap) {
  if (!h) return SVB_BAD_ARGUMENT;
  for (int i = 0; i < ST_COUNT && i < cap; ++i) {
    if (names) names[i] = kStageNames[i];
    if (ms) ms[i] = h->stage_ms[i];
  }
  return ST_COUNT;
}
int32_t svb_exchange_waits(SvbHandle* h, double* ms, int32_t cap, int32_t reset) {
  if (!h || h->multi) return SVB_BAD_ARGUMENT;
  CK(cudaSetDevice(h->device));
  if (int rc = sync_streams(h)) return rc;
  unsigned long long ns[8] = {};
  CK(cudaMemcpyFromSymbol(ns, g_wait_ns, sizeof ns));
  for (int i = 0; i < 6 && i < cap; ++i)
    if (ms) ms[i] = (double)ns[i] * 1e-6;
  if (ms && cap >= 8) { ms[6] = (double)h->n_ptiles; ms[7] = (double)h->n_tiles; }   // (as of the last front half the host has looked at)
  if (reset) {
    unsigned long long zero[8] = {};
    CK(cudaMemcpyToSymbol(g_wait_ns, zero, sizeof zero));
  }
  return cap >= 8 ? 8 : 6;
}
void svb_enable_stage_timing(SvbHandle* h, int32_t on) {
  if (h && h->multi) return;
  if (h) h->timing = on != 0;
}
void svb_set_option(SvbHandle* h, const char* name, double value) {
  if (!h || !name) return;
  if (h->multi) return multi_set_option(h, name, value);
  if (!std::strcmp(name, "store_grid")) h->store_grid = value != 0.0;  // CpuRunParameters::store_grid
  if (!std::strcmp(name, "global_particles")) h->n_global = (uint32_t)value;  // slab ranks: size of the original-order keyframe arrays
  if (!std::strcmp(name, "murmur_table_hash") && (value != 0.0) != h->murmur_hash) {
    // another hash function sends every key to another slot: start both front sets from empty tables
    h->murmur_hash = value != 0.0;
    for (auto& f : h->fs) f.fresh = true;
    h->binned_ahead = false;
  }
}

int32_t svb_node_ids_to_murmur(int32_t device, const int32_t* node_ids, const uint32_t* collider_bits, uint64_t n, uint32_t* hashes_node_ids, uint32_t* hashes_node_ids_and_bits) {
  if (!node_ids || n > 0x7fffffffull) return SVB_BAD_ARGUMENT;
  if (!n) return 0;
  SvbHandle* h = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SVB_CUDA_ERROR;  // no CPU fallback
  if (device < 0 || device >= count) return SVB_BAD_ARGUMENT;
  CK(cudaSetDevice(device));
  DevBuf ids, bits, out;
  CK(ids.ensure(n * 12));
  CK(bits.ensure(n * 4));
  CK(out.ensure(n * 8));
  CK(cudaMemcpy(ids.p, node_ids, n * 12, cudaMemcpyHostToDevice));
  if (collider_bits) CK(cudaMemcpy(bits.p, collider_bits, n * 4, cudaMemcpyHostToDevice));
  k_node_ids_to_murmur<<<blocks_for(n, 256), 256>>>(ids.as<int32_t>(), collider_bits ? bits.as<uint32_t>() : nullptr, (uint32_t)n, out.as<uint32_t>(), out.as<uint32_t>() + n);
  CK(cudaGetLastError());
  if (hashes_node_ids) CK(cudaMemcpy(hashes_node_ids, out.p, n * 4, cudaMemcpyDeviceToHost));
  if (hashes_node_ids_and_bits) CK(cudaMemcpy(hashes_node_ids_and_bits, out.as<uint32_t>() + n, n * 4, cudaMemcpyDeviceToHost));
  ids.release(); bits.release(); out.release();
  return 0;
}

int32_t svb_snapshot(SvbHandle* h) {
  if (!h) return SVB_BAD_ARGUMENT;
  if (h->multi) return fail(h, SVB_BAD_ARGUMENT, "svb_snapshot / svb_restore are single-device entry points");
  if (h->slabs) return fail(h, SVB_BAD_ARGUMENT, "svb_snapshot / svb_restore are single-device entry points (a slab rank's row count lives on the device)");
  if (int rc = set_device(h)) return rc;
  CK(h->snap_p.ensure(h->cap * NWORDS * 4));
  CK(h->snap_e.ensure(h->cap * 4));
  CK(cudaMemcpyAsync(h->snap_p.p, h->pbuf[h->cur].p, h->cap * NWORDS * 4, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->snap_e.p, h->energy.p, h->cap * 4, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->snap_time = h->time;
  h->snap_adaptive = h->adaptive;
  h->snap_substeps = h->substeps;
  h->snap_n = h->n;
  h->have_snapshot = true;
  return 0;
}
int32_t svb_restore(SvbHandle* h) {
  if (!h) return SVB_BAD_ARGUMENT;
  if (h->multi) return fail(h, SVB_BAD_ARGUMENT, "svb_snapshot / svb_restore are single-device entry points");
  if (h->slabs) return fail(h, SVB_BAD_ARGUMENT, "svb_snapshot / svb_restore are single-device entry points (a slab rank's row count lives on the device)");
  if (!h->have_snapshot) return fail(h, SVB_BAD_ARGUMENT, "svb_restore without svb_snapshot");
  if (int rc = set_device(h)) return rc;
  CK(cudaMemcpyAsync(h->pbuf[h->cur].p, h->snap_p.p, h->cap * NWORDS * 4, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->energy.p, h->snap_e.p, h->cap * 4, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->time = h->snap_time;
  h->adaptive = h->snap_adaptive;
  h->substeps = h->snap_substeps;
  h->n = h->snap_n;
  h->have_grid = false;
  h->binned_ahead = false;   // the bins in the other front set belong to the state that was just replaced
  h->limits_ahead = false;
  return 0;
}

}  // extern "C"

// ================================================================================================
// multi-GPU slab decomposition along x.  One process per GPU; neighbours exchange
// one block column of grid tiles after P2G and the particles that left the slab after the advance,
// with ncclSend / ncclRecv over NVLink.  Fixed time step only (the adaptive reductions would need an
// all-reduce per limit phase).
namespace {

#define NCK(call)                                                                                                   \
  do {                                                                                                              \
    ncclResult_t r_ = (call);                                                                                       \
    if (r_ != ncclSuccess) return fail(h, SVB_COMM_ERROR, "%s failed: %s (%s:%d)", #call, ncclGetErrorString(r_), __FILE__, __LINE__); \
  } while (0)

// re-stride the particle buffers to a larger per-quad capacity (quad q of row i = ((float4*)base)[q*cap + i])
int resize_particles(SvbHandle* h, size_t new_cap) {
  new_cap = (new_cap + 63) & ~(size_t)63;
  if (new_cap <= h->cap) return 0;
  CK(sync_streams(h));
  DevBuf nb[2], ne;
  for (int b = 0; b < 2; ++b) CK(nb[b].ensure(new_cap * NWORDS * 4));
  CK(ne.ensure(new_cap * 4));
  for (int q = 0; q < NQUADS; ++q)
    CK(cudaMemcpyAsync(nb[h->cur].as<uint32_t>() + (size_t)q * new_cap * 4, h->pbuf[h->cur].as<uint32_t>() + (size_t)q * h->cap * 4, h->cap * 16, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(ne.p, h->energy.p, h->cap * 4, cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int b = 0; b < 2; ++b) { h->pbuf[b].release(); h->pbuf[b] = nb[b]; }
  h->energy.release();
  h->energy = ne;
  h->cap = new_cap;
  CK(h->pcell.ensure(new_cap * 4));
  CK(h->prank.ensure(new_cap * 4));
  CK(h->src_of.ensure(new_cap * 4));
  h->have_snapshot = false;
  h->binned_ahead = false;   // pcell / prank were reallocated
  return 0;
}

// exchange two counters with the slab neighbours (left = rank-1, right = rank+1)
int exchange_counts(SvbHandle* h, const uint32_t send[2], uint32_t recv[2]) {
  uint32_t* d = h->comm_counts.as<uint32_t>();  // [0..1] send, [2..3] recv
  h->h_counts[0] = send[0]; h->h_counts[1] = send[1]; h->h_counts[2] = 0; h->h_counts[3] = 0;
  CK(cudaMemcpyAsync(d, h->h_counts, 16, cudaMemcpyHostToDevice, h->stream));
  NCK(ncclGroupStart());
  if (h->rank > 0) { NCK(ncclSend(d + 0, 1, ncclUint32, h->rank - 1, h->comm, h->stream)); NCK(ncclRecv(d + 2, 1, ncclUint32, h->rank - 1, h->comm, h->stream)); }
  if (h->rank + 1 < h->n_ranks) { NCK(ncclSend(d + 1, 1, ncclUint32, h->rank + 1, h->comm, h->stream)); NCK(ncclRecv(d + 3, 1, ncclUint32, h->rank + 1, h->comm, h->stream)); }
  NCK(ncclGroupEnd());
  CK(cudaMemcpyAsync(h->h_counts, d, 16, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  recv[0] = h->h_counts[2];
  recv[1] = h->h_counts[3];
  return 0;
}
int exchange_payload(SvbHandle* h, const void* send_l, const void* send_r, void* recv_l, void* recv_r, const uint32_t send[2], const uint32_t recv[2], size_t elem_bytes) {
  NCK(ncclGroupStart());
  if (h->rank > 0) {
    if (send[0]) NCK(ncclSend(send_l, (size_t)send[0] * elem_bytes, ncclUint8, h->rank - 1, h->comm, h->stream));
    if (recv[0]) NCK(ncclRecv(recv_l, (size_t)recv[0] * elem_bytes, ncclUint8, h->rank - 1, h->comm, h->stream));
  }
  if (h->rank + 1 < h->n_ranks) {
    if (send[1]) NCK(ncclSend(send_r, (size_t)send[1] * elem_bytes, ncclUint8, h->rank + 1, h->comm, h->stream));
    if (recv[1]) NCK(ncclRecv(recv_r, (size_t)recv[1] * elem_bytes, ncclUint8, h->rank + 1, h->comm, h->stream));
  }
  NCK(ncclGroupEnd());
  return 0;
}

// after P2G: add the neighbour's partial sums of the shared block column into this rank's tiles
int halo_exchange(SvbHandle* h) {
  cudaStream_t s = h->stream;
  StepScalars* S = cur_scalars(h);
  const TileTable T = tile_table(h);
  uint32_t* cnt = h->comm_counts.as<uint32_t>() + 4;  // [4] left, [5] right
  uint32_t sendc[2] = {0, 0}, recvc[2] = {0, 0};
  for (int attempt = 0;; ++attempt) {
    CK(cudaMemsetAsync(cnt, 0, 8, s));
    // my first column goes left (the left rank holds it as halo), my halo column (== hi) goes right
    if (h->rank > 0)
      k_pack_column<<<148 * 4, 256, 0, s>>>(S, T, h->layer_slots.as<unsigned long long>(), h->grid.as<float4>(), h->slab_lo, h->halo_send[0].as<HaloEntry>(), cnt + 0, (uint32_t)h->halo_cap);
    if (h->rank + 1 < h->n_ranks)
      k_pack_column<<<148 * 4, 256, 0, s>>>(S, T, h->layer_slots.as<unsigned long long>(), h->grid.as<float4>(), h->slab_hi, h->halo_send[1].as<HaloEntry>(), cnt + 1, (uint32_t)h->halo_cap);
    h->launches += 2;
    CK(cudaMemcpyAsync(h->h_counts + 4, cnt, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    sendc[0] = h->h_counts[4];
    sendc[1] = h->h_counts[5];
    if (int rc = exchange_counts(h, sendc, recvc)) return rc;
    const size_t need = std::max(std::max(sendc[0], sendc[1]), std::max(recvc[0], recvc[1]));
    if (need <= h->halo_cap) break;
    if (attempt > 2) return fail(h, SVB_COMM_ERROR, "halo buffers did not settle");
    const size_t c = need + need / 2 + 256;  // both ranks of a pair see the same counts, so both re-pack
    for (int k = 0; k < 2; ++k) { CK(h->halo_send[k].ensure(c * sizeof(HaloEntry))); CK(h->halo_recv[k].ensure(c * sizeof(HaloEntry))); }
    h->halo_cap = c;
  }
  if (int rc = exchange_payload(h, h->halo_send[0].p, h->halo_send[1].p, h->halo_recv[0].p, h->halo_recv[1].p, sendc, recvc, sizeof(HaloEntry))) return rc;
  // incoming tiles may be new to this rank: make room (tile arrays keep their contents)
  if ((size_t)h->n_tiles + recvc[0] + recvc[1] > h->tile_cap) return fail(h, SVB_COMM_ERROR, "tile capacity too small for the halo (%u + %u + %u > %zu)", h->n_tiles, recvc[0], recvc[1], h->tile_cap);
  for (int k = 0; k < 2; ++k)
    if (recvc[k]) {
      k_unpack_add<<<std::min<uint32_t>(blocks_for((uint64_t)recvc[k] * 32, 256), 148 * 8), 256, 0, s>>>(S, T, h->layer_slots.as<unsigned long long>(), h->layer_list.as<uint32_t>(),
                                                                                                         h->grid.as<float4>(), h->halo_recv[k].as<HaloEntry>(), recvc[k]);
      LAUNCH_CHECK();
    }
  h->halo_tiles_sent += sendc[0] + sendc[1];
  h->halo_margin = std::max<size_t>(h->halo_margin, 2 * ((size_t)recvc[0] + recvc[1]) + 2048);
  return 0;
}

// after the advance: hand particles that left [lo, hi) to the neighbour, append the ones coming in
int migrate(SvbHandle* h) {
  cudaStream_t s = h->stream;
  StepScalars* S = cur_scalars(h);
  uint32_t* cnt = h->comm_counts.as<uint32_t>() + 4;
  uint32_t sendc[2] = {0, 0}, recvc[2] = {0, 0};
  const uint32_t n = h->n;
  CK(cudaMemsetAsync(cnt, 0, 8, s));
  if (n) {
    k_migrate_pack<<<blocks_for(n, 256), 256, 0, s>>>(h->Pc(), h->energy.as<float>(), S, h->K.h, h->slab_lo, h->slab_hi, h->reach_lo, h->reach_hi, h->mig_send[0].as<uint32_t>(),
                                                     h->mig_send[1].as<uint32_t>(), cnt, (uint32_t)h->mig_cap, n);
    LAUNCH_CHECK();
  }
  CK(cudaMemcpyAsync(h->h_counts + 4, cnt, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h->h_scalars, S, sizeof(StepScalars), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (h->h_scalars->status & ST_KEY_RANGE) return fail(h, SVB_KEY_RANGE, "a particle crossed more than one slab in a single substep");
  sendc[0] = h->h_counts[4];
  sendc[1] = h->h_counts[5];
  if (std::max(sendc[0], sendc[1]) > h->mig_cap) return fail(h, SVB_COMM_ERROR, "migration buffer too small (%u rows > %zu)", std::max(sendc[0], sendc[1]), h->mig_cap);
  if (int rc = exchange_counts(h, sendc, recvc)) return rc;
  if (std::max(recvc[0], recvc[1]) > h->mig_cap) return fail(h, SVB_COMM_ERROR, "migration buffer too small (%u rows > %zu)", std::max(recvc[0], recvc[1]), h->mig_cap);
  if ((size_t)n + recvc[0] + recvc[1] > h->cap)
    if (int rc = resize_particles(h, ((size_t)n + recvc[0] + recvc[1]) * 5 / 4 + 4096)) return rc;
  if (int rc = exchange_payload(h, h->mig_send[0].p, h->mig_send[1].p, h->mig_recv[0].p, h->mig_recv[1].p, sendc, recvc, MIG_WORDS * 4)) return rc;
  uint32_t base = n;
  for (int k = 0; k < 2; ++k)
    if (recvc[k]) {
      k_migrate_unpack<<<blocks_for(recvc[k], 256), 256, 0, s>>>(h->Pc(), h->energy.as<float>(), h->mig_recv[k].as<uint32_t>(), recvc[k], base);
      LAUNCH_CHECK();
      base += recvc[k];
    }
  h->n = base;
  h->migrated_out += sendc[0] + sendc[1];
  return 0;
}


// this rank's mailbox: header | halo entries from the left | from the right | migrating rows from the left | from the right; sized
// from the largest slab so that every rank's mailbox has the same layout
int mailbox_alloc(SvbHandle* h, size_t n_max) {
  h->mb_halo_cap = n_max / 128 + 4096;
  h->mb_mig_cap = n_max / 4 + 65536;   // room for the columns a rebalance hands over at once, not just the per-substep trickle
  size_t off = 4096;  // header
  for (int k = 0; k < 2; ++k) { h->mb_halo_off[k] = off; off += h->mb_halo_cap * sizeof(HaloEntry); }
  for (int k = 0; k < 2; ++k) { h->mb_mig_off[k] = off; off += ((h->mb_mig_cap * MIG_WORDS * 4 + 255) & ~(size_t)255); }
  CK(h->mailbox.ensure(off));
  CK(cudaMemsetAsync(h->mailbox.p, 0, 4096, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}
// once every peer's mailbox is reachable (h->peer_mailbox): device-side scratch of the exchange kernels and the row count
int mailbox_finish(SvbHandle* h) {
  cudaStream_t s = h->stream;
  uint32_t* d = h->comm_counts.as<uint32_t>();
  h->p2p_local = d + 32;   // comm_counts: words 0..15 counts, 16.. slab table, 32..47 scratch of the sending kernels, 48 row count
  h->n_dev = d + 48;
  CK(cudaMemsetAsync(h->p2p_local, 0, 16 * 4, s));
  CK(cudaMemcpyAsync(h->n_dev, &h->n, 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemsetAsync(h->n_dev + 1, 0, 4, s));   // blocks-done tick of the rebalance receive
  CK(cudaStreamSynchronize(s));
  h->slab_seq = 0;
  h->dt_exchanges = 0;
  return 0;
}

// Peer-memory mailboxes: every rank allocates one buffer, shares it with CUDA IPC (handles travel through an NCCL
// all-gather), and maps every other rank's.  All ranks agree (all-reduce) on whether the mapping worked; if not,
// the NCCL send/recv path above carries the exchanges.  SVB_SLAB_NCCL=1 forces that path.
int setup_peer_mailboxes(SvbHandle* h) {
  const char* force = std::getenv("SVB_SLAB_NCCL");
  int ok = (force && force[0] == '1') || h->n_ranks > SLAB_MAX_RANKS ? 0 : 1;
  cudaStream_t s = h->stream;
  uint32_t* d = h->comm_counts.as<uint32_t>();
  // capacities from the largest slab, so that every mailbox has the same layout
  h->h_counts[0] = h->n;
  CK(cudaMemcpyAsync(d, h->h_counts, 4, cudaMemcpyHostToDevice, s));
  NCK(ncclAllReduce(d, d, 1, ncclUint32, ncclMax, h->comm, s));
  CK(cudaMemcpyAsync(h->h_counts, d, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const size_t n_max = h->h_counts[0];
  if (int rc = mailbox_alloc(h, n_max)) return rc;
  DevBuf handles;
  CK(handles.ensure((size_t)h->n_ranks * sizeof(cudaIpcMemHandle_t)));
  cudaIpcMemHandle_t mine;
  if (ok && cudaIpcGetMemHandle(&mine, h->mailbox.p) != cudaSuccess) { cudaGetLastError(); ok = 0; std::memset(&mine, 0, sizeof mine); }
  CK(cudaMemcpyAsync(handles.as<unsigned char>() + (size_t)h->rank * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice, s));
  NCK(ncclAllGather(handles.as<unsigned char>() + (size_t)h->rank * sizeof mine, handles.p, sizeof mine, ncclUint8, h->comm, s));
  std::vector<cudaIpcMemHandle_t> all(h->n_ranks);
  CK(cudaMemcpyAsync(all.data(), handles.p, all.size() * sizeof mine, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int r = 0; r < h->n_ranks && ok; ++r)
    if (r != h->rank && cudaIpcOpenMemHandle(&h->peer_mailbox[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      h->peer_mailbox[r] = nullptr;
      ok = 0;
    }
  // unanimous or not at all (this all-reduce also orders every rank's mailbox memset before the first message)
  h->h_counts[0] = (uint32_t)ok;
  CK(cudaMemcpyAsync(d, h->h_counts, 4, cudaMemcpyHostToDevice, s));
  NCK(ncclAllReduce(d, d, 1, ncclUint32, ncclMin, h->comm, s));
  CK(cudaMemcpyAsync(h->h_counts, d, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  handles.release();
  h->p2p = h->h_counts[0] != 0;
  h->ipc_mapped = h->p2p;
  if (!h->p2p) return 0;
  return mailbox_finish(h);
}

}  // namespace

extern "C" {

int32_t svb_comm_unique_id(uint8_t out[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return SVB_COMM_ERROR;
  std::memcpy(out, &id, 128);
  return 0;
}

int32_t svb_comm_init(SvbHandle* h, const uint8_t unique_id[128], int32_t rank, int32_t n_ranks, int32_t slab_lo_block_x, int32_t slab_hi_block_x, uint64_t original_offset) {
  if (!h || !unique_id || rank < 0 || rank >= n_ranks || slab_lo_block_x >= slab_hi_block_x) return SVB_BAD_ARGUMENT;
  if (int rc = set_device(h)) return rc;
  ncclUniqueId id;
  std::memcpy(&id, unique_id, 128);
  NCK(ncclCommInitRank(&h->comm, n_ranks, id, rank));
  h->rank = rank;
  h->n_ranks = n_ranks;
  h->slab_lo = slab_lo_block_x;
  h->slab_hi = slab_hi_block_x;
  // every rank learns every slab: a particle may only travel into the adjacent slab within one substep
  CK(h->comm_counts.ensure(256 + 64 + (size_t)n_ranks * 8));   // words 0..15 counts, 16.. slab table (2 per rank, after word 64), 32..47 p2p scratch, 48 row count
  CK(cudaMallocHost(&h->h_counts, 64 + (size_t)n_ranks * 8));
  int32_t* d_all = reinterpret_cast<int32_t*>(h->comm_counts.as<uint32_t>() + 16);
  int32_t mine[2] = {slab_lo_block_x, slab_hi_block_x};
  CK(cudaMemcpyAsync(d_all + 2 * rank, mine, 8, cudaMemcpyHostToDevice, h->stream));
  NCK(ncclAllGather(d_all + 2 * rank, d_all, 2, ncclInt32, h->comm, h->stream));
  std::vector<int32_t> all(2 * (size_t)n_ranks);
  CK(cudaMemcpyAsync(all.data(), d_all, all.size() * 4, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int r = 0; r + 1 < n_ranks; ++r)
    if (all[2 * r + 1] != all[2 * (r + 1)]) return fail(h, SVB_BAD_ARGUMENT, "slabs are not contiguous: rank %d ends at %d, rank %d starts at %d", r, all[2 * r + 1], r + 1, all[2 * (r + 1)]);
  h->reach_lo = rank > 0 ? all[2 * (rank - 1)] : slab_lo_block_x;
  h->reach_hi = rank + 1 < n_ranks ? all[2 * (rank + 1) + 1] : slab_hi_block_x;
  // global original indices travel with the particles
  if (h->n && original_offset) {
    k_add_orig<<<blocks_for(h->n, 256), 256, 0, h->stream>>>(h->Pc(), h->n, (uint32_t)original_offset);
    LAUNCH_CHECK();
  }
  h->orig_offset = original_offset;
  // room for incoming particles and halo tiles
  if (int rc = resize_particles(h, (size_t)h->n * 3 / 2 + 65536)) return rc;
  h->mig_cap = (size_t)h->n / 8 + 65536;
  h->halo_cap = 4096;
  for (int k = 0; k < 2; ++k) {
    CK(h->mig_send[k].ensure(h->mig_cap * MIG_WORDS * 4));
    CK(h->mig_recv[k].ensure(h->mig_cap * MIG_WORDS * 4));
    CK(h->halo_send[k].ensure(h->halo_cap * sizeof(HaloEntry)));
    CK(h->halo_recv[k].ensure(h->halo_cap * sizeof(HaloEntry)));
  }
  if (int rc = ensure_tile_capacity(h, h->tile_cap * 2 + 8192)) return rc;
  CK(cudaStreamSynchronize(h->stream));
  h->slabs = true;
  return setup_peer_mailboxes(h);
}

int32_t svb_slab_histogram(SvbHandle* h, int32_t first_col, uint32_t n_cols, uint64_t* counts) {
  if (!h || !counts) return SVB_BAD_ARGUMENT;
  if (int rc = set_device(h)) return rc;
  cudaStream_t s = h->stream;
  CK(cudaStreamSynchronize(s));
  if (h->p2p) CK(cudaMemcpy(&h->n, h->n_dev, 4, cudaMemcpyDeviceToHost));
  const size_t bytes = ((size_t)n_cols + 2) * 8;
  CK(h->node_offset.ensure(bytes));   // free between substeps
  CK(cudaMemsetAsync(h->node_offset.p, 0, bytes, s));
  if (h->n) {
    k_column_histogram<<<148 * 4, 256, 0, s>>>(h->Pc(), h->K.h, h->n, first_col, n_cols, h->node_offset.as<unsigned long long>());
    LAUNCH_CHECK();
  }
  CK(cudaMemcpyAsync(counts, h->node_offset.p, bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

int32_t svb_slab_rebalance(SvbHandle* h, int32_t new_lo, int32_t new_hi) {
  if (!h || new_lo >= new_hi) return SVB_BAD_ARGUMENT;
  if (!h->slabs || !h->p2p) return fail(h, SVB_COMM_ERROR, "svb_slab_rebalance needs a slab rank on the peer-memory path");
  if (int rc = set_device(h)) return rc;
  cudaStream_t s = h->stream;
  // every rank learns every old and new slab; the all-gather also orders this rebalance after every rank's last substep
  CK(h->scratch.ensure((size_t)h->n_ranks * 16 + 64));
  int32_t* d_tab = h->scratch.as<int32_t>();
  const int32_t mine[4] = {h->slab_lo, h->slab_hi, new_lo, new_hi};
  CK(cudaMemcpyAsync(d_tab + 4 * h->rank, mine, 16, cudaMemcpyHostToDevice, s));
  NCK(ncclAllGather(d_tab + 4 * h->rank, d_tab, 4, ncclInt32, h->comm, s));
  std::vector<int32_t> all(4 * (size_t)h->n_ranks);
  CK(cudaMemcpyAsync(all.data(), d_tab, all.size() * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int r = 0; r + 1 < h->n_ranks; ++r) {
    if (all[4 * r + 3] != all[4 * (r + 1) + 2]) return fail(h, SVB_BAD_ARGUMENT, "new slabs are not contiguous between ranks %d and %d", r, r + 1);
    // what a rank holds may only go to an adjacent rank: the new cut stays strictly inside the two old slabs it separates
    if (all[4 * r + 3] <= all[4 * r] || all[4 * r + 3] >= all[4 * (r + 1) + 1]) return fail(h, SVB_BAD_ARGUMENT, "the cut between ranks %d and %d moves beyond an adjacent slab", r, r + 1);
  }
  if (all[2] != all[0] || all[4 * (h->n_ranks - 1) + 3] != all[4 * (h->n_ranks - 1) + 1]) return fail(h, SVB_BAD_ARGUMENT, "the outer slab ends must stay where they are");
  const int reach_lo = h->rank > 0 ? all[4 * (h->rank - 1) + 2] : new_lo;
  const int reach_hi = h->rank + 1 < h->n_ranks ? all[4 * (h->rank + 1) + 3] : new_hi;
  StepScalars* S = cur_scalars(h);
  CK(h->mig_list.ensure(2 * h->mb_mig_cap * 4));
  CK(cudaMemcpy(&h->n, h->n_dev, 4, cudaMemcpyDeviceToHost));
  if (int rc = resize_particles(h, (size_t)h->n + 2 * h->mb_mig_cap)) return rc;   // a rebalance may hand over whole block columns
  const MigrateCut cut{new_lo, new_hi, reach_lo, reach_hi, 4.f * (float)new_lo, 4.f * (float)new_hi, h->mig_list.as<uint32_t>(), h->p2p_local + 8, (uint32_t)h->mb_mig_cap};
  const uint32_t seq = ++h->slab_seq;
  k_note_outside<<<148 * 4, 256, 0, s>>>(h->Pc(), S, h->K.h, cut, h->n_dev);
  LAUNCH_CHECK();
  SlabHeader* my_hdr = h->mailbox.as<SlabHeader>();
  unsigned char* my_mb = h->mailbox.as<unsigned char>();
  const bool has[2] = {h->rank > 0, h->rank + 1 < h->n_ranks};
  SlabPeers peers{};
  peers.n_ranks = h->n_ranks;
  for (int side = 0; side < 2; ++side)
    if (has[side]) {
      unsigned char* peer = static_cast<unsigned char*>(h->peer_mailbox[h->rank + (side ? 1 : -1)]);
      SlabHeader* ph = reinterpret_cast<SlabHeader*>(peer);
      const int their = side ? 0 : 1;
      peers.rows[side] = reinterpret_cast<uint32_t*>(peer + h->mb_mig_off[their]);
      peers.count[side] = &ph->mig_count[their];
      peers.seq[side] = &ph->mig_seq[their];
    }
  for (int r = 0; r < h->n_ranks; ++r)
    if (r != h->rank) {
      SlabHeader* ph = reinterpret_cast<SlabHeader*>(h->peer_mailbox[r]);
      peers.err_seq[r] = &ph->err_seq[h->rank];
      peers.err_val[r] = &ph->err_val[h->rank];
    }
  k_migrate_send_list<<<148, 256, 0, s>>>(h->Pc(), h->energy.as<float>(), S, cut, peers, (uint32_t)h->mb_mig_cap, seq, h->p2p_local + 10, 1, nullptr, nullptr, 1u);
  LAUNCH_CHECK();
  k_migrate_recv<<<148, 256, 0, s>>>(h->Pc(), h->energy.as<float>(), S, my_hdr, reinterpret_cast<const uint32_t*>(my_mb + h->mb_mig_off[0]), reinterpret_cast<const uint32_t*>(my_mb + h->mb_mig_off[1]),
                                     has[0] ? 1 : 0, has[1] ? 1 : 0, h->rank, h->n_ranks, seq, h->n_dev, /*between_substeps=*/1, h->K, BinNext{}, 0, peers);
  h->binned_ahead = false;   // rows left and arrived: the bins made ahead describe another row set
  LAUNCH_CHECK();
  if (int rc = read_status(h)) return rc;
  if (h->h_scalars->status & ST_COMM_OVERFLOW) return fail(h, SVB_COMM_ERROR, "rebalance: more particles change rank than the mailboxes (%zu rows) or the particle buffer hold", h->mb_mig_cap);
  if (h->h_scalars->status & ST_COMM_TIMEOUT) return fail(h, SVB_COMM_ERROR, "rebalance: a neighbour did not answer");
  if (h->h_scalars->status & ST_KEY_RANGE) return fail(h, SVB_KEY_RANGE, "rebalance: a particle would have to cross more than one slab");
  CK(cudaMemcpy(&h->n, h->n_dev, 4, cudaMemcpyDeviceToHost));
  h->slab_lo = new_lo;
  h->slab_hi = new_hi;
  h->reach_lo = reach_lo;
  h->reach_hi = reach_hi;
  return 0;
}

int32_t svb_set_original_indices(SvbHandle* h, const uint32_t* original_index, uint64_t n) {
  if (!h || !original_index || n != h->n) return SVB_BAD_ARGUMENT;
  if (int rc = set_device(h)) return rc;
  if (n) {
    uint32_t* st = h->pbuf[h->cur ^ 1].as<uint32_t>();   // the spare buffer is free between substeps
    CK(cudaMemcpyAsync(st, original_index, n * 4, cudaMemcpyHostToDevice, h->stream));
    k_wire_to_soa<1><<<blocks_for((uint32_t)n, 256), 256, 0, h->stream>>>(st, h->Pc(), PORIG, (uint32_t)n);
    LAUNCH_CHECK();
  }
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

int32_t svb_download_resident(SvbHandle* h, SvbParticles* out, uint64_t* original_index) {
  if (!h || !out) return SVB_BAD_ARGUMENT;
  if (int rc = set_device(h)) return rc;
  const uint32_t n = h->n;
  out->n = 0;
  if (!n) return 0;
  cudaStream_t s = h->stream;
  ParticleBuf P = h->Pc();
  uint32_t* rows = h->src_of.as<uint32_t>();  // free between substeps
  uint32_t* cnt = h->scratch.as<uint32_t>();
  CK(cudaMemsetAsync(cnt, 0, 4, s));
  k_resident_rows<<<blocks_for(n, 256), 256, 0, s>>>(P, n, rows, cnt);
  LAUNCH_CHECK();
  uint32_t m = 0;
  CK(cudaMemcpyAsync(&m, cnt, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  out->n = m;
  if (!m) return 0;
  float* stagef = h->pbuf[h->cur ^ 1].as<float>();
  const uint32_t blocks = blocks_for(m, 256);
  size_t stage_off = 0;
  auto field = [&](void* dst, int word, int k) -> int {
    if (!dst) return 0;
    float* st = stagef + stage_off;
    stage_off += (size_t)h->cap * k;
    if (k == 1) k_rows_to_wire<1><<<blocks, 256, 0, s>>>(P, word, rows, reinterpret_cast<uint32_t*>(st), m);
    else if (k == 3) k_rows_to_wire<3><<<blocks, 256, 0, s>>>(P, word, rows, reinterpret_cast<uint32_t*>(st), m);
    else k_rows_to_wire<9><<<blocks, 256, 0, s>>>(P, word, rows, reinterpret_cast<uint32_t*>(st), m);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(dst, st, (size_t)m * k * 4, cudaMemcpyDeviceToHost, s));
    return 0;
  };
  int rc = 0;
  std::vector<uint32_t> orig32(original_index ? m : 0);
  if ((rc = field(out->flags, PFLAGS, 1)) || (rc = field(out->mass, PMASS, 1)) || (rc = field(out->initial_volume, PVOL, 1)) ||
      (rc = field(out->mu_or_bulk_modulus, PP0, 1)) || (rc = field(out->lambda_or_exponent, PP1, 1)) || (rc = field(out->sand_alpha, PALPHA, 1)) ||
      (rc = field(out->viscosity_dynamic, PVD, 1)) || (rc = field(out->viscosity_bulk, PVB, 1)) || (rc = field(out->collider_bits, PBITS, 1)) ||
      (rc = field(out->positions, PX, 3)) || (rc = field(out->velocities, PV, 3)) || (rc = field(out->velocity_gradients, PC, 9)) ||
      (rc = field(out->position_gradients, PF, 9)) || (rc = field(original_index ? orig32.data() : nullptr, PORIG, 1)))
    return rc;
  if (out->elastic_energies) {
    CK(h->node_offset.ensure((size_t)m * 4));
    k_array_rows_to_wire<<<blocks, 256, 0, s>>>(h->energy.as<float>(), rows, h->node_offset.as<float>(), m);
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(out->elastic_energies, h->node_offset.p, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  if (original_index)
    for (uint32_t i = 0; i < m; ++i) original_index[i] = orig32[i];
  return 0;
}

}  // extern "C"

#include "svb_multi.inl"
