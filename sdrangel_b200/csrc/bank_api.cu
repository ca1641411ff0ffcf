// bank_api.cu — C ABI of the DownChannelizer bank (K3) and the per-channel plugin front-end (K4): b200dsp_bank_*
//
// Host-side mirror of (paths relative to the reference tree):
//   DownChannelizer::applyConfiguration / createFilterChain   sdrbase/dsp/downchannelizer.cpp:165-189,250-287
//   DownChannelizer::feed (output count / phase carry)        sdrbase/dsp/downchannelizer.cpp:50-91
//   NCO::setFreq                                              sdrbase/dsp/nco.cpp:48-52
//   Interpolator::create / createPolyphaseLowPass             sdrbase/dsp/interpolator.cpp:21-129
//   the NFM plugin's front-end wiring                         plugins/channelrx/demodnfm/nfmdemod.cpp:453-476
#include "common.cuh"
#include "hb48_tree.cuh"
#include "hb48_fused.cuh"
#include "hb48_chain.cuh"
#include "frontend.cuh"
#include <math.h>
#include <stdlib.h>
#include <string>
#include <vector>

using namespace b200dsp;

namespace {

// ---------------------------------------------------------------------------------------------------------
// small kernels around the level kernel
// ---------------------------------------------------------------------------------------------------------
struct GatherSrc { const void* src; const int* state; long long count; };     // state != null: count = state[3] (front-end)

// channel output = trunc_toward_zero(y / 2^S) per component (downchannelizer.cpp:78-83)
__global__ void hb48_finalize_kernel(const LeafChan* chans, const PassInfo pi)
{
    const LeafChan c = chans[blockIdx.y];
    if (c.depth == 0) return;                  // stage-less channel: forwarded by a plain copy
    if (c.direct && pi.fused_pass) return;     // written by the fused tree kernel itself
    const int n = pi.n_new[c.depth];
    const uint32_t* src = c.src + pi.wo[c.depth];
    uint32_t* dst = c.dst + pi.out_count[c.depth];
    const int bias = (1 << c.shift) - 1;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t w = src[k];
        int re = (int) (short) (w & 0xffffu), im = (int) w >> 16;
        re = (re + ((re >> 31) & bias)) >> c.shift;
        im = (im + ((im >> 31) & bias)) >> c.shift;
        dst[k] = ((uint32_t) re & 0xffffu) | ((uint32_t) im << 16);
    }
}

// pooled fetch: channel c's outputs of the last feed -> pool[c * stride ...], its count -> counts[c]
template <typename T>
__global__ void gather_outputs_kernel(const GatherSrc* srcs, T* pool, long long stride, long long* counts)
{
    const GatherSrc g = srcs[blockIdx.y];
    long long n = g.state ? (long long) g.state[3] : g.count;
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[blockIdx.y] = n;
    if (n > stride) n = stride;                 // reported through counts; the host call turns it into an error
    const T* src = static_cast<const T*>(g.src);
    T* dst = pool + (long long) blockIdx.y * stride;
    for (long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long) gridDim.x * blockDim.x) dst[k] = src[k];
}

__global__ void copy_words_kernel(const uint32_t* src, uint32_t* dst, long long n)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------------
// createFilterChain in the reference's float32/double mix (downchannelizer.cpp:250-287)
// ---------------------------------------------------------------------------------------------------------
bool band_contains(float ss, float se, float cs, float ce)
{
    if (se <= ss || ce <= cs) return false;
    return ss <= cs && se >= ce;
}

float filter_chain(float ss, float se, float cs, float ce, std::vector<int>& modes)
{
    while (modes.size() < 30) {
        const float bw = se - ss;
        const float quarter = bw / 4;
        const float low_end = (float) ((double) ss + (double) bw / 2.0);
        const float up_start = se - bw / 2.0f;
        if (band_contains(ss, low_end, cs, ce)) { modes.push_back(1); se = low_end; continue; }
        if (band_contains(up_start, se, cs, ce)) { modes.push_back(2); ss = up_start; continue; }
        const float cs2 = ss + quarter, ce2 = se - quarter;
        if (band_contains(cs2, ce2, cs, ce)) { modes.push_back(0); ss = cs2; se = ce2; continue; }
        break;
    }
    return (float) (((double) (float) (ce - cs) / 2.0 + (double) cs) - ((double) (float) (se - ss) / 2.0 + (double) ss));
}

struct Node { int depth, mode, parent, child[3], index; };      // index = position within its level

struct Channel {
    int requested_rate, center_offset;
    std::vector<int> modes;
    int node, S, out_rate, residual;
    int out_shift;          // the channel output is trunc(y / 2^out_shift): S for a whole chain (downchannelizer.cpp:78-83)
    // front-end
    bool fe;
    float nco_freq; int phase_steps; double cutoff, taps_per_phase; int fe_out_rate;
    int inc, ntaps; float ratio;
    std::vector<float> taps;
    // device
    uint32_t* d_out; long long out_cap; long long out_count;       // channelizer outputs since the last feed start
    uint32_t* d_hist; float* d_taps; float2* d_fe_out; int* d_sched; int* d_tile; int* d_state; long long* d_plan; long long fe_cap;
    long long A; int lattice, phshift;
    int scan_kb; long long scan_A;       // non-lattice ratio: parameters of the parallel exact replay
    int fe_lead;                         // the channel whose schedule this one shares (itself when it leads), see build()
};

const double PI_D = 3.14159265358979323846;

// Interpolator::create(phaseSteps, sampleRate, cutoff, nbTapsPerPhase): interpolator.cpp:21-56,74-110
void interp_taps(int phase_steps, double rate, double cutoff, double taps_per_phase, std::vector<float>& out, int* per_phase)
{
    int ntaps = (int) (taps_per_phase * phase_steps);
    if (ntaps % 2) ntaps++;
    const int np = ntaps;
    ntaps *= phase_steps;
    std::vector<float> taps(ntaps, 0.0f), window(ntaps);
    for (int n = 0; n < ntaps; n++) window[n] = (float) (0.54 - 0.46 * cos((2 * PI_D * n) / (ntaps - 1)));
    const int M = (ntaps - 1) / 2;
    const double fwT0 = 2 * PI_D * cutoff / (phase_steps * rate);
    for (int n = -M; n <= M; n++)
        taps[n + M] = (n == 0) ? (float) (fwT0 / PI_D * window[n + M]) : (float) (sin(n * fwT0) / (n * PI_D) * window[n + M]);
    double mx = taps[M];
    for (int n = 1; n <= M; n++) mx += 2.0 * taps[n + M];
    const double gain = 1.0 / mx;
    for (int i = 0; i < ntaps; i++) taps[i] = (float) (taps[i] * gain);
    out.assign(ntaps, 0.0f);
    for (int p = 0; p < phase_steps; p++) {
        float sum = 0;
        for (int i = 0; i < np; i++) { out[p * np + i] = taps[i * phase_steps + p]; sum += out[p * np + i]; }
        for (int i = 0; i < np; i++) out[p * np + i] /= sum;
    }
    *per_phase = np;
}

// non-lattice ratios in [1, 64): units and ratio for the warp-parallel exact replay (frontend.cuh: fe_scan_chunk)
bool scan_params(float ratio, int* kb_out, long long* A_out)
{
    *kb_out = 0; *A_out = 0;
    if (!(ratio > 1.0f && ratio < 64.0f)) return false;
    int e = 0;
    while ((float) (2 << e) <= ratio) ++e;               // 2^e <= ratio < 2^(e+1)
    const float P = (float) (2 << e);
    if (!(ratio + 1.0f > P)) return false;                // no power of two inside [ratio, ratio + 1): the lattice case
    const int kb = 23 - e;
    const double a = (double) ratio * (double) (1ll << kb);
    if (a != floor(a)) return false;
    *kb_out = kb; *A_out = (long long) a;
    return true;
}

// closed-form schedule when no sum r + ratio (< ratio + 1) can ever be rounded: ratio * 2^23 on the coarsest ulp lattice
bool lattice_params(float ratio, int phase_steps, long long* A_out, int* phshift_out)
{
    *A_out = 0; *phshift_out = 0;
    if (!(ratio >= 1.0f && ratio < 64.0f) || (phase_steps & (phase_steps - 1)) != 0) return false;
    const long long A = (long long) ((double) ratio * 8388608.0);
    long long smax = (long long) (((double) ratio + 1.0) * 8388608.0);
    int bl = 0;
    while (smax) { ++bl; smax >>= 1; }
    const long long g = 1ll << (bl > 24 ? bl - 24 : 0);
    int l2 = 0;
    while ((1 << l2) < phase_steps) ++l2;
    if ((double) A != (double) ratio * 8388608.0 || A % g != 0) return false;
    *A_out = A; *phshift_out = 23 - l2;
    return true;
}

} // namespace

struct b200dsp_interp {
    int device = 0; cudaStream_t stream = nullptr;
    int phase_steps = 0, ntaps = 0, parity = 0;
    std::vector<float> taps;
    float* d_taps = nullptr; uint32_t* d_hist = nullptr; int* d_state = nullptr; long long* d_plan = nullptr; FrontendChan* d_chan = nullptr;
    float2* d_in = nullptr; float2* d_out = nullptr; int* d_sched = nullptr; int* d_tile = nullptr; long long cap = 0, cap_out = 0;
};

struct b200dsp_bank {
    int device, sm_count;
    int input_rate;
    cudaStream_t stream, side;
    cudaEvent_t ev_begin, ev_sched;
    cudaStream_t copy; cudaEvent_t ev_h2d[2], ev_done[2];     // host-pointer feeds: H2D of pass p+1 under the kernels of pass p
    uint32_t* d_root_alt; long long root_alt_cap;              // second root staging buffer
    cudaEvent_t ev_tree0, ev_tree1; int tree_launches;         // timing of the tree-level launches of the last pass (b200dsp_bank_tree_time)
    std::vector<Channel> chans;
    bool built;
    // tree
    std::vector<Node> nodes;
    std::vector<std::vector<int>> levels;        // node ids per depth (levels[0] = {root})
    std::vector<std::vector<int>> fams;          // per depth d>=1: flat [parent index, child C, child L, child U] of families producing depth d
    int depth;
    // device
    long long chunk;                             // max input samples per internal pass
    std::vector<uint32_t*> d_level;  std::vector<long long> stride;   // per depth >= 1
    std::vector<uint32_t*> d_tail[2];            // per depth (parents), ping-pong
    std::vector<int*> d_fam;
    // fused multi-level launches (hb48_fused.cuh): the tree cut into depth ranges [b, b+k]
    struct FusedLaunch { int b, k, T; size_t smem; int n_groups, n_fams; FusedGroup* d_groups; FusedFam* d_fams; int lvl_off[FZ_MAXK]; };
    std::vector<FusedLaunch> flaunch;
    int chain_L; int chain_rot[CH_MAXL + 1]; int chain_index[CH_MAXL + 1];     // leading run of single-child levels (hb48_chain.cuh): depth, stage modes, node index per level
    bool fused_on;                               // B200DSP_NO_FUSED_TREE unset: aligned passes take hb48_fused_kernel
    // per-channel device buffers are rows of a few slabs (one allocation each, not seven per channel)
    uint32_t* slab_out; long long out_pitch;     // [channel][out_pitch] packed int16 IQ: channelizer outputs of the current feed
    float2* slab_fe; int* slab_sched; int* slab_tile; long long fe_pitch, tile_pitch;     // [channel][...] front-end outputs / schedule / tile table
    float* slab_taps; long long taps_pitch; int* slab_state; long long* slab_plan; uint32_t* slab_hist;      // [front-end][...] static + carried state
    cudaStream_t d2h; cudaEvent_t ev_pass;       // b200dsp_bank_process: device-to-host copies of finished output columns under the next pass
    std::vector<int> node_chan;                  // per node: the one channel ending there (-1 none, -2 several)
    std::vector<char> chan_direct;               // per channel: the fused kernel writes its output itself
    int n_indirect;                              // channels with stages that still need hb48_finalize_kernel after a fused pass
    uint32_t* d_root; long long root_cap;        // staging for host feeds / odd-pending device feeds
    std::vector<long long> produced;             // P[d]
    int tcur;
    LeafChan* d_leaf; FrontendChan* d_fe; float* d_nco;
    std::vector<LeafChan> h_leaf; std::vector<FrontendChan> h_fe;
    FrontendChan* d_fe_lead; std::vector<FrontendChan> h_fe_lead; std::vector<int> fe_leaders;    // the schedule kernel's table: one entry per distinct schedule
    int fe_parity;                               // current half of the front-ends' ping-pong carried state
    bool tables_dirty;                           // channel pointer tables must be re-uploaded (after a (re)allocation)
    std::vector<long long> out_count_depth;      // channel outputs per depth produced in the current feed
    std::vector<int> fe_index;                   // channel ids with a front-end
    // pooled fetch (b200dsp_bank_fetch_all): [channel][stride] staging + per-channel source table and counts
    void* d_pool; size_t pool_bytes;
    GatherSrc* d_gsrc; long long* d_gcnt; size_t gcap; int gslot; std::vector<GatherSrc> h_gsrc;
};

namespace {

void free_device(b200dsp_bank* b)
{
    for (auto p : b->d_level) if (p) cudaFree(p);
    for (int k = 0; k < 2; ++k) { for (auto p : b->d_tail[k]) if (p) cudaFree(p); b->d_tail[k].clear(); }
    for (auto p : b->d_fam) if (p) cudaFree(p);
    for (auto& fl : b->flaunch) { if (fl.d_groups) cudaFree(fl.d_groups); if (fl.d_fams) cudaFree(fl.d_fams); }
    b->flaunch.clear();
    b->d_level.clear(); b->d_fam.clear(); b->stride.clear();
    if (b->d_leaf) cudaFree(b->d_leaf);
    if (b->d_fe) cudaFree(b->d_fe);
    if (b->d_fe_lead) cudaFree(b->d_fe_lead);
    b->d_leaf = nullptr; b->d_fe = nullptr; b->d_fe_lead = nullptr;
    if (b->d_pool) cudaFree(b->d_pool);
    if (b->d_gsrc) cudaFree(b->d_gsrc);
    if (b->d_gcnt) cudaFree(b->d_gcnt);
    b->d_pool = nullptr; b->pool_bytes = 0; b->d_gsrc = nullptr; b->d_gcnt = nullptr; b->gcap = 0;
    void* slabs[] = { b->slab_out, b->slab_fe, b->slab_sched, b->slab_tile, b->slab_taps, b->slab_state, b->slab_plan, b->slab_hist };
    for (void* p : slabs) if (p) cudaFree(p);
    b->slab_out = nullptr; b->slab_fe = nullptr; b->slab_sched = nullptr; b->slab_tile = nullptr;
    b->slab_taps = nullptr; b->slab_state = nullptr; b->slab_plan = nullptr; b->slab_hist = nullptr;
    b->out_pitch = b->fe_pitch = b->tile_pitch = b->taps_pitch = 0;
    for (auto& c : b->chans) {
        c.d_out = nullptr; c.d_hist = nullptr; c.d_taps = nullptr; c.d_fe_out = nullptr; c.d_sched = nullptr; c.d_tile = nullptr; c.d_state = nullptr; c.d_plan = nullptr;
        c.out_cap = c.fe_cap = 0;
    }
    b->built = false;
}

// ---------------------------------------------------------------------------------------------------------
// Fused plan: cut the tree into depth ranges [b, b+k], k <= FZ_MAXK; every node of depth b that has children roots a group
// (hb48_fused.cuh).  k shrinks until the shared-memory pyramid of the widest group fits twice per SM.
// ---------------------------------------------------------------------------------------------------------
constexpr size_t FZ_SMEM_LIMIT = (size_t) (216 / FZ_CTAS_PER_SM) * 1024;      // the pyramid of a CTA: FZ_CTAS_PER_SM of them share an SM
inline int fused_tile0(int k) { const int a = HB_IN << (k - 1), b = HB_IN * FZ_WARPS; return a > b ? a : b; }   // smallest tile: whole items at the deepest level, one item per warp at the top

struct FusedBuild { std::vector<FusedGroup> groups; std::vector<FusedFam> fams; int max_slots[FZ_MAXK]; int min_items; };

void fused_collect(const b200dsp_bank* b, const std::vector<char>& is_leaf, int b0, int k, FusedBuild& out)
{
    out.groups.clear(); out.fams.clear();
    for (int j = 0; j < FZ_MAXK; ++j) out.max_slots[j] = 0;
    out.min_items = 1 << 30;
    const int T0 = fused_tile0(k);
    for (int rid : b->levels[b0]) {
        const Node& r = b->nodes[rid];
        if (r.child[0] < 0 && r.child[1] < 0 && r.child[2] < 0) continue;
        FusedGroup g;
        memset(&g, 0, sizeof(g));
        g.root_index = r.index; g.k = 0;
        std::vector<int> cur(1, rid), next;
        for (int j = 0; j < k; ++j) {
            g.fam_begin[j] = (int) out.fams.size();
            g.nslots[j] = (int) cur.size();
            next.clear();
            int nfam = 0;
            for (size_t s = 0; s < cur.size(); ++s) {
                const Node& n = b->nodes[cur[s]];
                if (n.child[0] < 0 && n.child[1] < 0 && n.child[2] < 0) continue;
                FusedFam f;
                f.pslot = (int) s; f.pindex = n.index; f.pbase = (int) s;       // pbase / cbase: slot numbers here, scaled by fused_scale_offsets
                for (int m = 0; m < 3; ++m) {
                    f.cbase[m] = -1; f.cbit[m] = 0; f.cindex[m] = -1;
                    const int c = n.child[m];
                    if (c < 0) continue;
                    const Node& cn = b->nodes[c];
                    const bool kids = cn.child[0] >= 0 || cn.child[1] >= 0 || cn.child[2] >= 0;
                    if (kids && j + 1 < k) { f.cbase[m] = (int) next.size(); f.cbit[m] = 1 << next.size(); next.push_back(c); }
                    if (kids && j + 1 == k) f.cindex[m] = cn.index;              // another launch reads the raw stage output
                    else if (is_leaf[c]) f.cindex[m] = (b->node_chan[c] >= 0) ? -2 - b->node_chan[c] : cn.index;
                }
                out.fams.push_back(f);
                ++nfam;
            }
            if (nfam) {
                g.k = j + 1;
                const int items = nfam * ((T0 >> j) / HB_IN);
                if (items < out.min_items) out.min_items = items;
            }
            cur.swap(next);
            if (cur.empty()) { for (int jj = j + 1; jj <= k; ++jj) g.fam_begin[jj] = (int) out.fams.size(); break; }
        }
        g.fam_begin[k] = (int) out.fams.size();
        for (int jj = g.k + 1; jj <= FZ_MAXK; ++jj) g.fam_begin[jj] = (int) out.fams.size();
        for (int j = 0; j < g.k; ++j) if (g.nslots[j] > out.max_slots[j]) out.max_slots[j] = g.nslots[j];
        out.groups.push_back(g);
    }
}

// slot numbers -> word offsets within the pyramid levels, for the launch's tile size
void fused_scale_offsets(FusedBuild& fb, int T)
{
    for (const FusedGroup& g : fb.groups)
        for (int j = 0; j < g.k; ++j)
            for (int f = g.fam_begin[j]; f < g.fam_begin[j + 1]; ++f) {
                FusedFam& x = fb.fams[f];
                x.pbase = x.pslot * 4 * (HB_HIST + (T >> (j + 1)));
                for (int m = 0; m < 3; ++m) if (x.cbase[m] >= 0) x.cbase[m] *= 4 * (HB_HIST + (T >> (j + 2)));
            }
}

size_t fused_smem(const FusedBuild& fb, int k, int T, int* lvl_off)
{
    size_t words = 0;
    for (int j = 0; j < k; ++j) {
        if (lvl_off) lvl_off[j] = (int) words;
        words += (size_t) fb.max_slots[j] * 4 * (HB_HIST + (T >> (j + 1)));
    }
    return words * 4;
}

int build_fused_plan(b200dsp_bank* b, const std::vector<char>& is_leaf)
{
    int rc;
    const int D = b->depth;
    // a node where exactly one channel ends, and which no later launch reads, is finalised by the fused kernel itself
    b->node_chan.assign(b->nodes.size(), -1);
    for (size_t i = 0; i < b->chans.size(); ++i) {
        int& nc = b->node_chan[b->chans[i].node];
        nc = (nc == -1) ? (int) i : -2;
    }
    b->chan_direct.assign(b->chans.size(), 0);
    // leading chain: while the node has exactly one child, and that child is not itself a channel's last stage
    b->chain_L = 0;
    b->chain_index[0] = 0;
    for (int cur = 0; b->chain_L < CH_MAXL && b->chain_L < D - 1;) {
        const Node& nd = b->nodes[cur];
        int kids = 0, m1 = -1;
        for (int m = 0; m < 3; ++m) if (nd.child[m] >= 0) { ++kids; m1 = m; }
        if (kids != 1 || is_leaf[nd.child[m1]]) break;
        cur = nd.child[m1];
        ++b->chain_L;
        b->chain_rot[b->chain_L] = (m1 == 0) ? 0 : (m1 == 1 ? 1 : -1);
        b->chain_index[b->chain_L] = b->nodes[cur].index;
    }
    int b0 = b->chain_L;
    while (b0 < D) {
        const int left = D - b0;
        const int parts = (left + FZ_MAXK - 1) / FZ_MAXK;
        int k = (left + parts - 1) / parts;
        FusedBuild fb;
        for (;; --k) {
            fused_collect(b, is_leaf, b0, k, fb);
            bool wide = false;
            for (int j = 0; j < k; ++j) if (fb.max_slots[j] > 32) wide = true;
            if ((!wide && fused_smem(fb, k, fused_tile0(k), nullptr) <= FZ_SMEM_LIMIT) || k == 1) break;
        }
        b200dsp_bank::FusedLaunch fl;
        memset(&fl, 0, sizeof(fl));
        fl.b = b0; fl.k = k; fl.T = fused_tile0(k);
        // sparse groups (chains) have few items per level at the smallest tile: a larger tile keeps the warps busy
        while (fl.T < 12288 && fb.min_items * (fl.T / fused_tile0(k)) < FZ_WARPS && fused_smem(fb, k, 2 * fl.T, nullptr) <= FZ_SMEM_LIMIT) fl.T *= 2;
        fl.smem = fused_smem(fb, k, fl.T, fl.lvl_off);
        fused_scale_offsets(fb, fl.T);
        fl.n_groups = (int) fb.groups.size(); fl.n_fams = (int) fb.fams.size();
        if (fl.smem > 200 * 1024) { b->flaunch.clear(); b->chan_direct.assign(b->chans.size(), 0); return 0; }        // a tree this wide stays on the one-level kernel
        if (fl.n_groups) {
            if ((rc = B200_CUDA_CHECK(cudaMalloc(&fl.d_groups, fb.groups.size() * sizeof(FusedGroup)))) ||
                (rc = B200_CUDA_CHECK(cudaMemcpy(fl.d_groups, fb.groups.data(), fb.groups.size() * sizeof(FusedGroup), cudaMemcpyHostToDevice))) ||
                (rc = B200_CUDA_CHECK(cudaMalloc(&fl.d_fams, fb.fams.size() * sizeof(FusedFam)))) ||
                (rc = B200_CUDA_CHECK(cudaMemcpy(fl.d_fams, fb.fams.data(), fb.fams.size() * sizeof(FusedFam), cudaMemcpyHostToDevice)))) return rc;
        }
        for (const FusedFam& f : fb.fams)
            for (int m = 0; m < 3; ++m) if (f.cindex[m] <= -2) b->chan_direct[-2 - f.cindex[m]] = 1;
        b->flaunch.push_back(fl);
        b0 += k;
    }
    return 0;
}

// Build the shared-prefix tree and all device state (filter histories zero, like freshly constructed reference objects).
int build(b200dsp_bank* b)
{
    free_device(b);
    b->nodes.clear(); b->levels.clear(); b->fams.clear();
    Node root = { 0, 0, -1, { -1, -1, -1 }, 0 };
    b->nodes.push_back(root);
    b->depth = 0;
    for (auto& c : b->chans) {
        int cur = 0;
        for (size_t s = 0; s < c.modes.size(); ++s) {
            const int m = c.modes[s];
            if (b->nodes[cur].child[m] < 0) {
                Node n = { (int) s + 1, m, cur, { -1, -1, -1 }, 0 };
                b->nodes[cur].child[m] = (int) b->nodes.size();
                b->nodes.push_back(n);
            }
            cur = b->nodes[cur].child[m];
        }
        c.node = cur;
        if ((int) c.modes.size() > b->depth) b->depth = (int) c.modes.size();
    }
    b->levels.assign(b->depth + 1, std::vector<int>());
    for (size_t i = 0; i < b->nodes.size(); ++i) {
        Node& n = b->nodes[i];
        n.index = (int) b->levels[n.depth].size();
        b->levels[n.depth].push_back((int) i);
    }
    b->fams.assign(b->depth + 1, std::vector<int>());
    for (int d = 1; d <= b->depth; ++d)
        for (int id : b->levels[d - 1]) {
            const Node& n = b->nodes[id];
            if (n.child[0] < 0 && n.child[1] < 0 && n.child[2] < 0) continue;
            b->fams[d].push_back(n.index);
            for (int m = 0; m < 3; ++m) b->fams[d].push_back(n.child[m] >= 0 ? b->nodes[n.child[m]].index : -1);
        }
    std::vector<char> is_leaf(b->nodes.size(), 0);
    for (auto& c : b->chans) is_leaf[c.node] = 1;
    int rc;
    b->d_level.assign(b->depth + 1, nullptr); b->stride.assign(b->depth + 1, 0);
    b->d_tail[0].assign(b->depth + 1, nullptr); b->d_tail[1].assign(b->depth + 1, nullptr);
    b->d_fam.assign(b->depth + 1, nullptr);
    for (int d = 1; d <= b->depth; ++d) {
        b->stride[d] = (((b->chunk >> d) + 16) + 3) & ~3ll;
        const size_t bytes = (size_t) b->levels[d].size() * b->stride[d] * 4;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_level[d], bytes))) || (rc = B200_CUDA_CHECK(cudaMemset(b->d_level[d], 0, bytes)))) return rc;
        const size_t fb = b->fams[d].size() * sizeof(int);
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_fam[d], fb))) ||
            (rc = B200_CUDA_CHECK(cudaMemcpy(b->d_fam[d], b->fams[d].data(), fb, cudaMemcpyHostToDevice)))) return rc;
    }
    if ((rc = build_fused_plan(b, is_leaf))) return rc;
    for (int d = 0; d < b->depth; ++d)
        for (int k = 0; k < 2; ++k) {
            const size_t bytes = (size_t) b->levels[d].size() * TAIL_WORDS * 4;
            if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_tail[k][d], bytes))) || (rc = B200_CUDA_CHECK(cudaMemset(b->d_tail[k][d], 0, bytes)))) return rc;
        }
    b->produced.assign(b->depth + 1, 0);
    b->tcur = 0;
    b->fe_parity = 0;
    const size_t nc = b->chans.size();
    if (nc) {
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_leaf, nc * sizeof(LeafChan)))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&b->d_fe, nc * sizeof(FrontendChan)))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&b->d_fe_lead, nc * sizeof(FrontendChan))))) return rc;
    }
    b->h_leaf.resize(nc);
    b->tables_dirty = true;
    b->out_count_depth.assign(32, 0);
    b->fe_index.clear();
    b->fe_leaders.clear();
    size_t max_taps = 0;
    for (size_t i = 0; i < nc; ++i) {
        Channel& c = b->chans[i];
        c.out_count = 0;
        if (!c.fe) continue;
        b->fe_index.push_back((int) i);
        if (c.taps.size() > max_taps) max_taps = c.taps.size();
    }
    if (const size_t nfe = b->fe_index.size()) {
        b->taps_pitch = (long long) ((max_taps + 3) & ~(size_t) 3);
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->slab_taps, nfe * b->taps_pitch * sizeof(float)))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&b->slab_state, nfe * 4 * sizeof(int)))) || (rc = B200_CUDA_CHECK(cudaMemset(b->slab_state, 0, nfe * 4 * sizeof(int)))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&b->slab_plan, nfe * 4 * sizeof(long long)))) || (rc = B200_CUDA_CHECK(cudaMemset(b->slab_plan, 0, nfe * 4 * sizeof(long long)))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&b->slab_hist, nfe * 2 * FE_HIST_WORDS * 4))) || (rc = B200_CUDA_CHECK(cudaMemset(b->slab_hist, 0, nfe * 2 * FE_HIST_WORDS * 4)))) return rc;
        std::vector<float> all(nfe * (size_t) b->taps_pitch, 0.0f);
        for (size_t k = 0; k < nfe; ++k) {
            Channel& c = b->chans[b->fe_index[k]];
            std::copy(c.taps.begin(), c.taps.end(), all.begin() + k * b->taps_pitch);
            c.d_taps = b->slab_taps + k * b->taps_pitch;
            c.d_hist = b->slab_hist + 2 * FE_HIST_WORDS * k;
            // The schedule (which samples emit an output, at which phase) depends only on the ratio, the channel's depth and the
            // distance carried so far -- not on the data.  All channels of a bank see the same feeds from the same zero state, so
            // channels with equal (depth, ratio, phase steps) have the same schedule for ever: one of them leads, the others share
            // its schedule, tile table, plan and counters (64 NFM channels at 10 MS/s: 2 schedules; the 1024-channel plan: 1).
            size_t lead = k;
            for (size_t q = 0; q < k; ++q) {
                const Channel& o = b->chans[b->fe_index[q]];
                if (o.fe_lead == b->fe_index[q] && o.S == c.S && memcmp(&o.ratio, &c.ratio, sizeof(float)) == 0 && o.phase_steps == c.phase_steps &&
                    o.lattice == c.lattice && o.A == c.A && o.phshift == c.phshift && o.scan_kb == c.scan_kb && o.scan_A == c.scan_A) { lead = q; break; }
            }
            c.fe_lead = b->fe_index[lead];
            if (lead == k) b->fe_leaders.push_back(b->fe_index[k]);
            c.d_state = b->slab_state + 4 * lead; c.d_plan = b->slab_plan + 4 * lead;
        }
        if ((rc = B200_CUDA_CHECK(cudaMemcpy(b->slab_taps, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice)))) return rc;
    }
    b->h_fe.resize(b->fe_index.size());
    if (!b->d_nco) {
        std::vector<float> t(4096);
        for (int i = 0; i < 4096; i++) t[i] = (float) cos((2.0 * PI_D * i) / 4096);       // NCO::initTable, nco.cpp:30-39
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_nco, 4096 * sizeof(float)))) ||
            (rc = B200_CUDA_CHECK(cudaMemcpy(b->d_nco, t.data(), 4096 * sizeof(float), cudaMemcpyHostToDevice)))) return rc;
    }
    // the allocations above were zeroed/filled on the legacy default stream; the bank works on non-blocking streams
    if ((rc = B200_CUDA_CHECK(cudaDeviceSynchronize()))) return rc;
    b->built = true;
    return 0;
}

// make sure the output slabs can take the outputs of a feed of n input samples (rows of one pitch: the longest channel's)
int reserve_outputs(b200dsp_bank* b, long long n)
{
    int rc;
    const size_t nc = b->chans.size();
    if (nc == 0) return 0;
    long long need = 0;
    for (auto& c : b->chans) if ((n >> c.S) + 2 > need) need = (n >> c.S) + 2;
    need = (need + 3) & ~3ll;
    if (b->out_pitch < need) {
        if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(b->stream)))) return rc;
        if (b->slab_out) cudaFree(b->slab_out);
        b->slab_out = nullptr; b->out_pitch = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->slab_out, nc * (size_t) need * 4)))) return rc;
        b->out_pitch = need;
        for (size_t i = 0; i < nc; ++i) { b->chans[i].d_out = b->slab_out + i * (size_t) need; b->chans[i].out_cap = need; }
        b->tables_dirty = true;
    }
    if (!b->fe_index.empty() && b->fe_pitch < need) {
        if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(b->stream)))) return rc;
        if (b->slab_fe) cudaFree(b->slab_fe);
        if (b->slab_sched) cudaFree(b->slab_sched);
        if (b->slab_tile) cudaFree(b->slab_tile);
        b->slab_fe = nullptr; b->slab_sched = nullptr; b->slab_tile = nullptr; b->fe_pitch = 0;
        const long long tp = need / FE_TILE + 4;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->slab_fe, nc * (size_t) need * sizeof(float2)))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&b->slab_sched, nc * (size_t) need * sizeof(int)))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&b->slab_tile, nc * (size_t) tp * sizeof(int))))) return rc;
        b->fe_pitch = need; b->tile_pitch = tp;
        for (size_t i = 0; i < nc; ++i) {
            Channel& c = b->chans[i];
            if (!c.fe) continue;
            c.d_fe_out = b->slab_fe + i * (size_t) need;
            c.d_sched = b->slab_sched + (size_t) c.fe_lead * (size_t) need; c.d_tile = b->slab_tile + (size_t) c.fe_lead * (size_t) tp;
            c.fe_cap = need;
        }
        b->tables_dirty = true;
    }
    return 0;
}

// one internal pass over n <= chunk input samples available at device pointer d_in (16-byte aligned)
int feed_chunk(b200dsp_bank* b, const uint32_t* d_in, long long n, cudaStream_t st, bool first_pass)
{
    int rc;
    const int D = b->depth;
    std::vector<long long> Pb = b->produced, Pa(D + 1);
    Pa[0] = Pb[0] + n;
    for (int d = 1; d <= D; ++d) Pa[d] = Pa[d - 1] / 2;
    // root buffer: B[i] = root sample (C_before + i); with one pending sample the new data has to sit at B[1]
    const uint32_t* rootB = d_in;
    const int root_pend = (D >= 1) ? (int) (Pb[0] - 2 * Pb[1]) : 0;
    if (root_pend || ((uintptr_t) d_in & 15)) {
        const long long need = n + 8;
        if (b->root_cap < need) {
            if (b->d_root) cudaFree(b->d_root);
            b->d_root = nullptr; b->root_cap = 0;
            if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_root, (size_t) need * 4)))) return rc;
            b->root_cap = need;
        }
        if (d_in != b->d_root + root_pend) {
            copy_words_kernel<<<(unsigned) ((n + 1023) / 1024 < 2048 ? (n + 1023) / 1024 : 2048), 256, 0, st>>>(d_in, b->d_root + root_pend, n);
            if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
        }
        rootB = b->d_root;
    }
    // channel outputs of this pass: static pointer tables (re-uploaded only after a reallocation) + per-depth scalars
    const size_t nc = b->chans.size();
    PassInfo pi;
    memset(&pi, 0, sizeof(pi));
    pi.first_pass = first_pass ? 1 : 0;
    pi.parity = b->fe_parity;
    int max_new = 0;
    for (int d = 0; d <= D && d < 32; ++d) {
        pi.n_new[d] = (int) (Pa[d] - Pb[d]);
        pi.wo[d] = (d == 0) ? 0 : (int) (Pb[d] & 1);
        pi.out_count[d] = b->out_count_depth[d];
    }
    for (size_t i = 0; i < nc; ++i) if (pi.n_new[b->chans[i].S] > max_new) max_new = pi.n_new[b->chans[i].S];
    if (b->tables_dirty && nc) {
        for (size_t i = 0; i < nc; ++i) {
            Channel& c = b->chans[i];
            b->h_leaf[i].src = (c.S == 0) ? nullptr : b->d_level[c.S] + (long long) b->nodes[c.node].index * b->stride[c.S];
            b->h_leaf[i].dst = c.d_out; b->h_leaf[i].depth = c.S; b->h_leaf[i].shift = c.out_shift;
            b->h_leaf[i].direct = (i < b->chan_direct.size() && b->chan_direct[i]) ? 1 : 0; b->h_leaf[i].pad = 0;
        }
        for (size_t k = 0; k < b->fe_index.size(); ++k) {
            Channel& c = b->chans[b->fe_index[k]];
            FrontendChan& f = b->h_fe[k];
            f.in = c.d_out; f.hist = c.d_hist; f.taps = c.d_taps; f.out = c.d_fe_out; f.sched = c.d_sched; f.tile_start = c.d_tile; f.state = c.d_state; f.plan = c.d_plan; f.A = c.scan_kb ? c.scan_A : c.A; f.scan_kb = c.scan_kb; f.lattice = c.lattice; f.phshift = c.phshift; f.in_f32 = 0; f.hist_stride = FE_HIST_WORDS;
            f.depth = c.S; f.inc = c.inc; f.ntaps = c.ntaps; f.phase_steps = c.phase_steps; f.ratio = c.ratio; f.mode = 0; f.sched_cap = (int) c.fe_cap;
        }
        if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(b->d_leaf, b->h_leaf.data(), nc * sizeof(LeafChan), cudaMemcpyHostToDevice, st)))) return rc;
        if (!b->h_fe.empty() && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(b->d_fe, b->h_fe.data(), b->h_fe.size() * sizeof(FrontendChan), cudaMemcpyHostToDevice, st)))) return rc;
        b->h_fe_lead.clear();
        for (size_t k = 0; k < b->fe_index.size(); ++k) if (b->chans[b->fe_index[k]].fe_lead == b->fe_index[k]) b->h_fe_lead.push_back(b->h_fe[k]);
        if (!b->h_fe_lead.empty() && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(b->d_fe_lead, b->h_fe_lead.data(), b->h_fe_lead.size() * sizeof(FrontendChan), cudaMemcpyHostToDevice, st)))) return rc;
        b->tables_dirty = false;
    }
    // front-end schedules depend only on counts: replay them on the side stream while the tree runs
    // (the first pass of a feed always runs it, even with nothing new at the channels' depths: it zeroes the per-feed output counts)
    const bool run_sched = !b->fe_index.empty() && (max_new > 0 || first_pass);
    if (run_sched) {
        if ((rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_begin, st))) || (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(b->side, b->ev_begin, 0)))) return rc;
        const int nfe = (int) b->fe_leaders.size();
        frontend_schedule_kernel<<<nfe, 32 * FE_SW, 0, b->side>>>(b->d_fe_lead, nfe, pi);      // one CTA per distinct schedule
        if ((rc = B200_CUDA_CHECK(cudaGetLastError())) || (rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_sched, b->side)))) return rc;
    }
    const int tc = b->tcur, tn = tc ^ 1;
    if ((rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_tree0, st)))) return rc;
    b->tree_launches = 0;
    // fused path: the pass starts aligned at every level (no pending sample, even pair index) and is a multiple of 2^depth long
    bool fused = b->fused_on && !b->flaunch.empty() && D >= 1 && n > 0 && (n % (1ll << D)) == 0 && n < (1ll << 31);
    for (int d = 1; d <= D && fused; ++d) fused = (Pb[d - 1] == 2 * Pb[d]) && !(Pb[d] & 1);
    if (fused && b->chain_L > 0) {
        // the leading single-child levels as one warp-private cascade (hb48_chain.cuh); the pyramid launches start below it
        const int L = b->chain_L;
        ChainParams q;
        memset(&q, 0, sizeof(q));
        q.in = rootB; q.out = b->d_level[L] + (long long) b->chain_index[L] * b->stride[L];
        for (int s2 = 0; s2 < L; ++s2) {
            q.tail_in[s2] = b->d_tail[tc][s2] + (long long) b->chain_index[s2] * TAIL_WORDS;
            q.tail_out[s2] = b->d_tail[tn][s2] + (long long) b->chain_index[s2] * TAIL_WORDS;
        }
        for (int s2 = 1; s2 <= L; ++s2) q.rot[s2] = (signed char) b->chain_rot[s2];
        q.n0 = n; q.L = L; q.opq_zero = 0; q.opq_one = 1; q.opq_mone = -1;
        const long long U = (long long) HB_IN << (L - 1);
        const long long sp_total = (n + U - 1) / U;
        const int wpb = (L <= 4) ? 8 : 4;
        const size_t smem = (size_t) wpb * L * HB_STAGE_BYTES;
        if ((rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) hb48_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)))) return rc;
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, (const void*) hb48_chain_kernel, wpb * 32, smem) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 1; }
        const long long max_warps = (long long) b->sm_count * nb * wpb;
        long long slice_sp = (sp_total + max_warps - 1) / max_warps;
        if (slice_sp < 4) slice_sp = (sp_total < 4) ? sp_total : 4;          // a slice pays one superphase of warm-up
        if (slice_sp < 1) slice_sp = 1;
        q.slice_sp = (int) slice_sp; q.n_slices = (int) ((sp_total + slice_sp - 1) / slice_sp);
        hb48_chain_kernel<<<(unsigned) ((q.n_slices + wpb - 1) / wpb), wpb * 32, smem, st>>>(q);
        if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
        ++b->tree_launches;
    }
    if (fused) {
        // A schedule kernel CTA on the scan path (256 threads x 128 registers, ~0.1 ms) takes half an SM's registers, i.e. one of
        // that SM's two pyramid CTA slots -- both when two of them land on one SM; the pyramid's ranges are a static partition,
        // so a CTA without a slot would run after the others.  The launch that overlaps the schedule kernel therefore gets two
        // CTAs fewer per schedule CTA, all of them resident at once (measured on the 64-channel plan, 2 schedule CTAs: 296 / 294
        // pyramid CTAs 0.56 ms per step, 292 and 288: 0.486).
        int displaced = 0;
        if (run_sched) for (int ci : b->fe_leaders) if (b->chans[ci].scan_kb > 0) displaced += 2;
        if (displaced > b->sm_count / 4) displaced = b->sm_count / 4;
        for (const auto& fl : b->flaunch) {
            if (fl.n_groups == 0) continue;
            FusedParams q;
            memset(&q, 0, sizeof(q));
            q.in_base = (fl.b == 0) ? rootB : b->d_level[fl.b];
            q.in_stride = (fl.b == 0) ? 0 : b->stride[fl.b];
            for (int j = 0; j < fl.k; ++j) {
                q.out_base[j] = b->d_level[fl.b + j + 1]; q.out_stride[j] = b->stride[fl.b + j + 1];
                q.tail_in[j] = b->d_tail[tc][fl.b + j]; q.tail_out[j] = b->d_tail[tn][fl.b + j];
                q.lvl_off[j] = fl.lvl_off[j];
            }
            q.groups = fl.d_groups; q.fams = fl.d_fams; q.n_groups = fl.n_groups;
            q.leaf = b->d_leaf;
            for (int j = 0; j < fl.k; ++j) q.leaf_count[j] = b->out_count_depth[fl.b + j + 1];
            q.lvl_off[fl.k] = 0;
            q.T = fl.T; q.l2items = 0;
            while ((HB_IN << q.l2items) < fl.T) ++q.l2items;
            q.n_root = (int) (n >> fl.b); q.tpr = (q.n_root + fl.T - 1) / fl.T;
            q.opq_zero = 0; q.opq_one = 1; q.opq_mone = -1;
            const long long tot = (long long) q.n_groups * q.tpr;
            // persistent CTAs over contiguous (group, tile) ranges; a range that starts inside a stream pays one warm-up tile
            long long ctas = (long long) b->sm_count * ((fl.smem <= FZ_SMEM_LIMIT) ? FZ_CTAS_PER_SM : 1);
            if (ctas > displaced + b->sm_count) ctas -= displaced;
            displaced = 0;                                        // later launches start after the schedule kernel has drained
            if (ctas > (tot + 3) / 4) ctas = (tot + 3) / 4;
            if (ctas < 1) ctas = 1;
            if ((rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) hb48_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) fl.smem)))) return rc;
            hb48_fused_kernel<<<(unsigned) ctas, FZ_THREADS, fl.smem, st>>>(q);
            if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
            ++b->tree_launches;
        }
    }
    for (int d = 1; d <= D && !fused; ++d) {
        const int n_fam = (int) (b->fams[d].size() / 4);
        if (n_fam == 0) continue;
        LevelParams p;
        memset(&p, 0, sizeof(p));
        p.in_base = (d == 1) ? rootB : b->d_level[d - 1];
        p.in_stride = (d == 1) ? 0 : b->stride[d - 1];
        p.out_base = b->d_level[d]; p.out_stride = b->stride[d];
        p.tail_in = b->d_tail[tc][d - 1]; p.tail_out = b->d_tail[tn][d - 1];
        p.fam = b->d_fam[d]; p.n_fam = n_fam;
        const long long Cb = 2 * Pb[d], Ca = 2 * Pa[d];
        p.n_in = (int) (Ca - Cb);
        p.pend = (int) (Pb[d - 1] - Cb);
        p.in_limit = p.pend + (int) (Pa[d - 1] - Pb[d - 1]);
        p.wo = (int) (Pb[d] & 1);
        p.flip = (int) (Pb[d] & 1);
        p.opq_zero = 0; p.opq_one = 1; p.opq_mone = -1;
        {
            const int nb = (p.n_in + HB_IN - 1) / HB_IN;
            // Many short slices (8 batches = 3072 samples each, 64 samples of history re-read per slice): the grid is several
            // waves deep, so losing SMs to a concurrent kernel (the NCCL broadcast of the next step) costs a few percent, not
            // a whole extra wave.  Small levels still get at least one slice per resident warp slot.
            long long target = (long long) b->sm_count * 16;
            long long slices = (target + n_fam - 1) / n_fam;
            const long long by8 = (nb + 7) / 8;
            if (slices < by8) slices = by8;
            if (slices > nb) slices = nb;
            if (slices < 1) slices = 1;
            const int bps = (int) ((nb + slices - 1) / slices);
            if (bps > 0) slices = (nb + bps - 1) / bps;
            p.slices = (int) slices; p.bps = bps;
            const long long warps = (long long) n_fam * slices;
            const int wpb = 4;
            long long blocks = (warps + wpb - 1) / wpb;
            hb48_level_kernel<<<(unsigned) blocks, wpb * 32, wpb * HB_STAGE_BYTES, st>>>(p);
            if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
            ++b->tree_launches;
        }
    }
    if ((rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_tree1, st)))) return rc;
    // stage-less channels (S == 0) are forwarded unchanged (downchannelizer.cpp:57-60): plain copy of the input
    for (size_t i = 0; i < nc; ++i) {
        Channel& c = b->chans[i];
        if (c.S == 0 && n > 0) {
            if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(c.d_out + b->out_count_depth[0], d_in, (size_t) n * 4, cudaMemcpyDeviceToDevice, st)))) return rc;
        }
    }
    pi.fused_pass = fused ? 1 : 0;
    bool need_finalize = !fused;
    if (fused) for (size_t i = 0; i < nc; ++i) if (b->chans[i].S > 0 && !b->chan_direct[i]) { need_finalize = true; break; }
    if (nc && max_new > 0 && D >= 1 && need_finalize) {
        int gx = (max_new + 255) / 256;
        if (gx > 64) gx = 64;
        hb48_finalize_kernel<<<dim3(gx, (unsigned) nc), 256, 0, st>>>(b->d_leaf, pi);
        if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
    }
    if (run_sched && (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(st, b->ev_sched, 0)))) return rc;
    if (!b->fe_index.empty() && max_new > 0) {
        size_t smem = 0;
        for (int ci : b->fe_index) { const Channel& cc = b->chans[ci]; const size_t s2 = (((size_t) ((cc.ntaps + 2 * FE_PAD) | 1) * cc.phase_steps + 3) & ~(size_t) 3) * sizeof(float); if (s2 > smem) smem = s2; }
        smem += (size_t) (FE_MAX_TAPS + FE_TILE + FE_Z_EXTRA) * sizeof(float2);
        // every channel a 5/4 closed-form resampler with 72 taps per phase (the 1024-channel plan): register-tiled variant
        bool lat54 = true;
        for (int ci : b->fe_index) { const Channel& cc = b->chans[ci]; if (!cc.lattice || 4 * cc.A != (5ll << 23) || cc.ntaps != 72) { lat54 = false; break; } }
        if (lat54) {
            // closed-form outputs per channel <= 0.8 * inputs + 1
            const long long max_out = (long long) max_new * 4 / 5 + 2;
            const dim3 grid((unsigned) ((max_out + F54_OPB - 1) / F54_OPB), (unsigned) b->h_fe.size());
            if ((rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) frontend54_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F54_SMEM)))) return rc;
            frontend54_kernel<<<grid, F54_THREADS, F54_SMEM, st>>>(b->d_fe, b->d_nco, pi);
        } else {
            const dim3 grid((unsigned) ((max_new + FE_TILE - 1) / FE_TILE), (unsigned) b->h_fe.size());
            if (smem > 48 * 1024 && (rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) frontend_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)))) return rc;
            frontend_kernel_t<false><<<grid, FE_THREADS, smem, st>>>(b->d_fe, b->d_nco, pi);
        }
        if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
        b->fe_parity ^= 1;
    }
    for (int d = 0; d <= D && d < 32; ++d) b->out_count_depth[d] += pi.n_new[d];
    for (size_t i = 0; i < nc; ++i) b->chans[i].out_count = b->out_count_depth[b->chans[i].S];
    b->produced = Pa;
    b->tcur = tn;
    return 0;
}

// length of the next internal pass: at most `chunk`; while the stream is aligned at every level, long passes are cut to a
// multiple of 2^(depth+1) so that they take the fused kernel and leave the stream aligned (a ragged rest follows as its
// own short pass through the one-level kernel)
long long next_pass_len(const b200dsp_bank* b, long long remaining)
{
    long long m = remaining < b->chunk ? remaining : b->chunk;
    const int D = b->depth;
    if (!b->fused_on || b->flaunch.empty() || D < 1 || D > 28) return m;
    const long long unit = 2ll << D;
    if (m < unit || m % unit == 0) return m;
    for (int d = 1; d <= D; ++d) if (b->produced[d - 1] != 2 * b->produced[d] || (b->produced[d] & 1)) return m;
    return m - m % unit;
}

// the two staging buffers of the host-pointer feeds (a pass is at most `chunk` samples)
int ensure_root_buffers(b200dsp_bank* b, long long n_samples)
{
    int rc;
    const long long first_m = n_samples < b->chunk ? n_samples : b->chunk;
    if (b->root_cap >= first_m + 8 && b->root_alt_cap >= first_m + 8) return 0;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(b->stream))) || (rc = B200_CUDA_CHECK(cudaStreamSynchronize(b->copy)))) return rc;
    const long long cap = (b->chunk > first_m ? b->chunk : first_m) + 8;
    if (b->root_cap < cap) {
        if (b->d_root) cudaFree(b->d_root);
        b->d_root = nullptr; b->root_cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_root, (size_t) cap * 4)))) return rc;
        b->root_cap = cap;
    }
    if (b->root_alt_cap < cap) {
        if (b->d_root_alt) cudaFree(b->d_root_alt);
        b->d_root_alt = nullptr; b->root_alt_cap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_root_alt, (size_t) cap * 4)))) return rc;
        b->root_alt_cap = cap;
    }
    return 0;
}

// a feed of zero samples (DownChannelizer::feed with begin == end): nothing moves, every channel's outputs of "the last
// feed" are empty -- the front-ends' per-feed counts live on the device and are zeroed by an empty schedule pass
int empty_feed(b200dsp_bank* b, cudaStream_t st)
{
    if (b->fe_index.empty()) return 0;
    PassInfo pi;
    memset(&pi, 0, sizeof(pi));
    pi.first_pass = 1;
    const int nfe = (int) b->fe_leaders.size();
    if (b->tables_dirty) return 0;            // never fed: the device tables do not exist yet and every count is still zero
    frontend_schedule_kernel<<<nfe, 32 * FE_SW, 0, st>>>(b->d_fe_lead, nfe, pi);
    return B200_CUDA_CHECK(cudaGetLastError());
}

int feed_common(b200dsp_bank* b, const uint32_t* d_in, long long n, cudaStream_t st)
{
    int rc;
    if (!b->built && (rc = build(b))) return rc;
    if ((rc = reserve_outputs(b, n))) return rc;
    for (auto& c : b->chans) c.out_count = 0;
    b->out_count_depth.assign(32, 0);
    if (n == 0) return empty_feed(b, st);
    long long done = 0;
    while (done < n) {
        const long long m = next_pass_len(b, n - done);
        if ((rc = feed_chunk(b, d_in + done, m, st, done == 0))) return rc;
        done += m;
    }
    return 0;
}

} // namespace

extern "C" {

int b200dsp_filter_chain(int input_rate_hz, int requested_rate_hz, int center_offset_hz, int* out_rate_hz, int* residual_offset_hz, int* modes, int cap)
{
    if (input_rate_hz <= 0 || requested_rate_hz <= 0) return b200_fail(B200DSP_EINVAL, "filter_chain: rates must be positive");
    std::vector<int> m;
    const float ofs = filter_chain((float) (input_rate_hz / -2), (float) (input_rate_hz / 2),
                                   (float) (center_offset_hz - requested_rate_hz / 2), (float) (center_offset_hz + requested_rate_hz / 2), m);
    if (out_rate_hz) *out_rate_hz = input_rate_hz / (1 << m.size());
    if (residual_offset_hz) *residual_offset_hz = (int) ofs;
    for (int i = 0; i < (int) m.size() && i < cap; ++i) if (modes) modes[i] = m[i];
    return (int) m.size();
}

int b200dsp_bank_create(b200dsp_bank_t** out, int input_rate_hz)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "bank_create: null handle pointer");
    *out = nullptr;
    if (input_rate_hz <= 0) return b200_fail(B200DSP_EINVAL, "bank_create: input rate must be positive");
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_bank* b = new (std::nothrow) b200dsp_bank();
    if (!b) return b200_fail(B200DSP_ENOMEM, "bank_create: out of host memory");
    b->device = b200_current_device();
    b->sm_count = b200_sm_count_of(b->device);
    b->input_rate = input_rate_hz;
    b->built = false; b->depth = 0; b->chunk = 3ll << 22; b->tables_dirty = true;
    b->fused_on = (getenv("B200DSP_NO_FUSED_TREE") == nullptr);
    b->d_root = nullptr; b->root_cap = 0;
    b->slab_out = nullptr; b->slab_fe = nullptr; b->slab_sched = nullptr; b->slab_tile = nullptr;
    b->slab_taps = nullptr; b->slab_state = nullptr; b->slab_plan = nullptr; b->slab_hist = nullptr;
    b->out_pitch = b->fe_pitch = b->tile_pitch = b->taps_pitch = 0; b->d2h = nullptr; b->ev_pass = nullptr;
    b->d_leaf = nullptr; b->d_fe = nullptr; b->d_fe_lead = nullptr; b->d_nco = nullptr; b->tcur = 0;
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(b->device))) || (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithPriority(&b->side, cudaStreamNonBlocking, -5))) ||
        (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_begin, cudaEventDisableTiming))) ||
        (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_sched, cudaEventDisableTiming))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&b->copy, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&b->d2h, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_pass, cudaEventDisableTiming)))) { delete b; return rc; }
    if ((rc = B200_CUDA_CHECK(cudaEventCreate(&b->ev_tree0))) || (rc = B200_CUDA_CHECK(cudaEventCreate(&b->ev_tree1)))) { delete b; return rc; }
    for (int i = 0; i < 2; ++i)
        if ((rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_h2d[i], cudaEventDisableTiming))) ||
            (rc = B200_CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_done[i], cudaEventDisableTiming)))) { delete b; return rc; }
    *out = b;
    return 0;
}

int b200dsp_bank_destroy(b200dsp_bank_t* b)
{
    if (!b) return 0;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    cudaStreamSynchronize(b->side);
    free_device(b);
    if (b->copy) cudaStreamSynchronize(b->copy);
    if (b->d_root) cudaFree(b->d_root);
    if (b->d_root_alt) cudaFree(b->d_root_alt);
    if (b->d_nco) cudaFree(b->d_nco);
    if (b->copy) cudaStreamDestroy(b->copy);
    if (b->d2h) { cudaStreamSynchronize(b->d2h); cudaStreamDestroy(b->d2h); }
    if (b->ev_pass) cudaEventDestroy(b->ev_pass);
    for (int i = 0; i < 2; ++i) { if (b->ev_h2d[i]) cudaEventDestroy(b->ev_h2d[i]); if (b->ev_done[i]) cudaEventDestroy(b->ev_done[i]); }
    cudaStreamDestroy(b->stream);
    cudaStreamDestroy(b->side);
    if (b->ev_tree0) cudaEventDestroy(b->ev_tree0);
    if (b->ev_tree1) cudaEventDestroy(b->ev_tree1);
    cudaEventDestroy(b->ev_begin);
    cudaEventDestroy(b->ev_sched);
    delete b;
    return 0;
}

int b200dsp_bank_set_chunk(b200dsp_bank_t* b, int64_t samples)
{
    if (!b || samples < 768 || samples > (1ll << 30)) return b200_fail(B200DSP_EINVAL, "bank_set_chunk: chunk must be within [768, 2^30] samples (kernels index a pass with 32-bit integers)");
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    const long long chunk = (samples + 767) / 768 * 768;
    for (const auto& c : b->chans)
        if (c.fe && !c.lattice && (chunk >> c.S) >= (1ll << 24)) return b200_fail(B200DSP_EINVAL, "bank_set_chunk: a channel's replayed schedule indexes a pass with 24 bits; chunk too large for it");
    b->chunk = chunk;
    free_device(b);
    return 0;
}

int b200dsp_bank_add_channel(b200dsp_bank_t* b, int requested_rate_hz, int center_offset_hz, int* chan_id, int* out_rate_hz, int* residual_offset_hz)
{
    if (!b) return b200_fail(B200DSP_EINVAL, "null handle");
    if (requested_rate_hz <= 0) return b200_fail(B200DSP_EINVAL, "bank_add_channel: requested rate must be positive");
    Channel c{};
    c.requested_rate = requested_rate_hz; c.center_offset = center_offset_hz;
    c.fe = false; c.d_out = nullptr; c.out_cap = 0; c.out_count = 0;
    c.d_hist = nullptr; c.d_taps = nullptr; c.d_fe_out = nullptr; c.d_sched = nullptr; c.d_state = nullptr; c.d_plan = nullptr; c.fe_cap = 0;
    // downchannelizer.cpp:169-171: integer divides first, then int -> Real
    const float ofs = filter_chain((float) (b->input_rate / -2), (float) (b->input_rate / 2),
                                   (float) (center_offset_hz - requested_rate_hz / 2), (float) (center_offset_hz + requested_rate_hz / 2), c.modes);
    c.S = (int) c.modes.size();
    c.out_shift = c.S;
    c.out_rate = b->input_rate / (1 << c.S);
    c.residual = (int) ofs;
    if (b->built) { cudaSetDevice(b->device); cudaStreamSynchronize(b->stream); free_device(b); }   // plan changes: state restarts
    b->chans.push_back(c);
    if (chan_id) *chan_id = (int) b->chans.size() - 1;
    if (out_rate_hz) *out_rate_hz = c.out_rate;
    if (residual_offset_hz) *residual_offset_hz = c.residual;
    return 0;
}

// A channel given by its filter stages instead of (rate, offset).  Used to split a bank across GPUs: a bank fed with the
// output of tree node P (depth k) holds the channels below P by their path suffix, with out_shift = k + len(suffix), and a
// "top" bank exposes the depth-k nodes themselves as channels with out_shift = 0 (raw stage output).
int b200dsp_bank_add_channel_path(b200dsp_bank_t* b, const int* modes, int n_modes, int out_shift, int* chan_id)
{
    if (!b || n_modes < 0 || n_modes > 30 || (n_modes > 0 && !modes) || out_shift < 0 || out_shift > 30) return b200_fail(B200DSP_EINVAL, "bank_add_channel_path: bad argument");
    Channel c{};
    for (int i = 0; i < n_modes; ++i) {
        if (modes[i] < 0 || modes[i] > 2) return b200_fail(B200DSP_EINVAL, "bank_add_channel_path: stage mode must be 0, 1 or 2");
        c.modes.push_back(modes[i]);
    }
    c.S = n_modes; c.out_shift = out_shift;
    c.out_rate = b->input_rate / (1 << c.S);
    if (b->built) { cudaSetDevice(b->device); cudaStreamSynchronize(b->stream); free_device(b); }
    b->chans.push_back(c);
    if (chan_id) *chan_id = (int) b->chans.size() - 1;
    return 0;
}

// back to the state of freshly constructed reference objects (zero filter histories, phase 0) without rebuilding the plan
int b200dsp_bank_reset(b200dsp_bank_t* b, void* cuda_stream)
{
    if (!b) return b200_fail(B200DSP_EINVAL, "null handle");
    if (!b->built) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : b->stream;
    for (int d = 0; d < b->depth; ++d)
        for (int k = 0; k < 2; ++k)
            if ((rc = B200_CUDA_CHECK(cudaMemsetAsync(b->d_tail[k][d], 0, (size_t) b->levels[d].size() * TAIL_WORDS * 4, st)))) return rc;
    for (auto& c : b->chans) c.out_count = 0;
    if (const size_t nfe = b->fe_index.size()) {
        if ((rc = B200_CUDA_CHECK(cudaMemsetAsync(b->slab_state, 0, nfe * 4 * sizeof(int), st))) ||
            (rc = B200_CUDA_CHECK(cudaMemsetAsync(b->slab_hist, 0, nfe * 2 * FE_HIST_WORDS * 4, st))) ||
            (rc = B200_CUDA_CHECK(cudaMemsetAsync(b->slab_plan, 0, nfe * 4 * sizeof(long long), st)))) return rc;
    }
    b->produced.assign(b->depth + 1, 0);
    b->out_count_depth.assign(32, 0);
    b->tcur = 0; b->fe_parity = 0;
    return 0;
}

// Device time of the tree-level launches (hb48_level_kernel, one per level) of the last internal pass, between two events on
// the feed's stream: bench.py's per-launch roofline of the dominant kernel.  Waits for that pass to finish.
int b200dsp_bank_tree_time(b200dsp_bank_t* b, float* ms, int* launches)
{
    if (!b || !ms) return b200_fail(B200DSP_EINVAL, "bank_tree_time: bad argument");
    if (!b->built || b->tree_launches == 0) return b200_fail(B200DSP_ESTATE, "bank_tree_time: no feed with tree levels yet");
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaEventSynchronize(b->ev_tree1))) || (rc = B200_CUDA_CHECK(cudaEventElapsedTime(ms, b->ev_tree0, b->ev_tree1)))) return rc;
    if (launches) *launches = b->tree_launches;
    return 0;
}

int b200dsp_bank_channel_path(b200dsp_bank_t* b, int chan_id, int* modes, int cap)
{
    if (!b || chan_id < 0 || chan_id >= (int) b->chans.size()) return b200_fail(B200DSP_EINVAL, "bank_channel_path: bad channel");
    const Channel& c = b->chans[chan_id];
    for (int i = 0; i < c.S && i < cap; ++i) if (modes) modes[i] = c.modes[i];
    return c.S;
}

int b200dsp_bank_node_count(b200dsp_bank_t* b)
{
    if (!b) return b200_fail(B200DSP_EINVAL, "null handle");
    if (!b->built) { int rc = B200_CUDA_CHECK(cudaSetDevice(b->device)); if (rc) return rc; if ((rc = build(b))) return rc; }
    return (int) b->nodes.size() - 1;
}

int b200dsp_bank_set_frontend(b200dsp_bank_t* b, int chan_id, float nco_freq_hz, int phase_steps, double cutoff_hz, double taps_per_phase, int out_rate_hz)
{
    if (!b || chan_id < 0 || chan_id >= (int) b->chans.size()) return b200_fail(B200DSP_EINVAL, "bank_set_frontend: bad channel");
    if (phase_steps < 1 || phase_steps > 255 || out_rate_hz <= 0 || taps_per_phase <= 0) return b200_fail(B200DSP_EINVAL, "bank_set_frontend: bad parameters");
    Channel& c = b->chans[chan_id];
    // computed aside and committed only once everything is validated: a failed call leaves the channel as it was
    int np = 0;
    std::vector<float> taps;
    interp_taps(phase_steps, (double) c.out_rate, cutoff_hz, taps_per_phase, taps, &np);
    if (np > FE_MAX_TAPS) return b200_fail(B200DSP_EINVAL, "bank_set_frontend: %d taps per phase exceed the supported %d", np, FE_MAX_TAPS);
    const size_t smem = ((((size_t) ((np + 2 * FE_PAD) | 1) * phase_steps + 3) & ~(size_t) 3)) * sizeof(float) + (size_t) (FE_MAX_TAPS + FE_TILE + FE_Z_EXTRA) * sizeof(float2);
    if (smem > 200 * 1024) return b200_fail(B200DSP_EINVAL, "bank_set_frontend: %d phases x %d taps need %zu bytes of shared memory (limit 200 KiB)", phase_steps, np, smem);
    const float ratio = (float) c.out_rate / (float) out_rate_hz;                  // nfmdemod.cpp:469-470
    long long A = 0; int phshift = 0;
    const int lattice = lattice_params(ratio, phase_steps, &A, &phshift) ? 1 : 0;
    // the replayed schedule packs the input index of an output into 24 bits (frontend.cuh): a pass must stay below 2^24 channel samples
    if (!lattice && (b->chunk >> c.S) >= (1ll << 24))
        return b200_fail(B200DSP_EINVAL, "bank_set_frontend: a pass of %lld samples at this channel's rate exceeds 2^24; lower the chunk (b200dsp_bank_set_chunk)", (long long) (b->chunk >> c.S));
    c.taps.swap(taps);
    c.fe = true; c.nco_freq = nco_freq_hz; c.phase_steps = phase_steps; c.cutoff = cutoff_hz; c.taps_per_phase = taps_per_phase; c.fe_out_rate = out_rate_hz;
    c.ntaps = np;
    c.inc = (int) ((nco_freq_hz * 4096) / (float) c.out_rate);                 // NCO::setFreq: float arithmetic, truncation (nco.cpp:50)
    c.ratio = ratio; c.lattice = lattice; c.A = A; c.phshift = phshift;
    c.scan_kb = 0; c.scan_A = 0;
    if (!lattice && !getenv("B200DSP_SERIAL_SCHEDULE")) scan_params(ratio, &c.scan_kb, &c.scan_A);
    if (b->built) { cudaSetDevice(b->device); cudaStreamSynchronize(b->stream); free_device(b); }
    return 0;
}

int b200dsp_bank_frontend_info(b200dsp_bank_t* b, int chan_id, int* nco_increment, int* taps_per_phase, float* taps, int taps_cap)
{
    if (!b || chan_id < 0 || chan_id >= (int) b->chans.size() || !b->chans[chan_id].fe) return b200_fail(B200DSP_EINVAL, "bank_frontend_info: no front-end on this channel");
    const Channel& c = b->chans[chan_id];
    if (nco_increment) *nco_increment = c.inc;
    if (taps_per_phase) *taps_per_phase = c.ntaps;
    if (taps) for (int i = 0; i < (int) c.taps.size() && i < taps_cap; ++i) taps[i] = c.taps[i];
    return 0;
}

int b200dsp_bank_feed_dev(b200dsp_bank_t* b, const void* d_iq, int64_t n_samples, void* cuda_stream)
{
    if (!b) return b200_fail(B200DSP_EINVAL, "null handle");
    if (n_samples < 0 || (n_samples > 0 && !d_iq)) return b200_fail(B200DSP_EINVAL, "bank_feed: bad buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    return feed_common(b, (const uint32_t*) d_iq, n_samples, cuda_stream ? (cudaStream_t) cuda_stream : b->stream);
}

int b200dsp_bank_feed(b200dsp_bank_t* b, const int16_t* iq, int64_t n_samples)
{
    if (!b) return b200_fail(B200DSP_EINVAL, "null handle");
    if (n_samples < 0 || (n_samples > 0 && !iq)) return b200_fail(B200DSP_EINVAL, "bank_feed: bad buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    if (!b->built && (rc = build(b))) return rc;
    if ((rc = reserve_outputs(b, n_samples))) return rc;
    for (auto& c : b->chans) c.out_count = 0;
    b->out_count_depth.assign(32, 0);
    if (n_samples == 0) return empty_feed(b, b->stream);
    // Passes of `chunk` samples through two staging buffers: the H2D copy of pass p+1 (copy stream) runs under the kernels of
    // pass p (bank stream); a buffer is refilled only after the pass that read it has finished (ev_done).
    if ((rc = ensure_root_buffers(b, n_samples))) return rc;
    long long done = 0;
    int pass = 0;
    while (done < n_samples) {
        const long long m = next_pass_len(b, n_samples - done);
        const int pend = (b->depth >= 1) ? (int) (b->produced[0] - 2 * b->produced[1]) : 0;
        const int slot = pass & 1;
        // ev_done[slot] is only waited on once pass-2 has recorded it in this call; an earlier call's passes were synchronised at its end
        if (pass >= 2 && (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(b->copy, b->ev_done[slot], 0)))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(b->d_root + pend, iq + 2 * done, (size_t) m * 4, cudaMemcpyHostToDevice, b->copy))) ||
            (rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_h2d[slot], b->copy))) ||
            (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(b->stream, b->ev_h2d[slot], 0)))) return rc;
        if ((rc = feed_chunk(b, b->d_root + pend, m, b->stream, done == 0))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_done[slot], b->stream)))) return rc;
        std::swap(b->d_root, b->d_root_alt);          // the next pass stages into the other buffer
        std::swap(b->root_cap, b->root_alt_cap);
        done += m;
        ++pass;
    }
    return B200_CUDA_CHECK(cudaStreamSynchronize(b->stream));
}

int b200dsp_bank_fetch(b200dsp_bank_t* b, int chan_id, int stage, void* out, int64_t cap, int64_t* n)
{
    if (!b || chan_id < 0 || chan_id >= (int) b->chans.size()) return b200_fail(B200DSP_EINVAL, "bank_fetch: bad channel");
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(b->stream)))) return rc;
    Channel& c = b->chans[chan_id];
    if (stage == B200DSP_STAGE_CHANNELIZER) {
        if (n) *n = c.out_count;
        if (c.out_count > cap) return b200_fail(B200DSP_EINVAL, "bank_fetch: buffer too small (%lld needed)", c.out_count);
        if (c.out_count && out) return B200_CUDA_CHECK(cudaMemcpy(out, c.d_out, (size_t) c.out_count * 4, cudaMemcpyDeviceToHost));
        return 0;
    }
    if (stage == B200DSP_STAGE_FRONTEND) {
        if (!c.fe) return b200_fail(B200DSP_ESTATE, "bank_fetch: channel has no front-end");
        long long total = 0;
        if (c.d_state && b->built && c.d_fe_out) {
            int s[4];
            if ((rc = B200_CUDA_CHECK(cudaMemcpy(s, c.d_state, sizeof(s), cudaMemcpyDeviceToHost)))) return rc;
            total = s[3];
        }
        if (n) *n = total;
        if (total > cap) return b200_fail(B200DSP_EINVAL, "bank_fetch: buffer too small (%lld needed)", total);
        if (total && out) return B200_CUDA_CHECK(cudaMemcpy(out, c.d_fe_out, (size_t) total * sizeof(float2), cudaMemcpyDeviceToHost));
        return 0;
    }
    return b200_fail(B200DSP_EINVAL, "bank_fetch: bad stage");
}

// The front-end schedule of the channel's LAST internal pass: for each output the index of the channel sample that emitted
// it (within that pass) and the polyphase filter phase -- what Interpolator::decimate's distance recurrence decided.  Listed
// outputs come from the schedule kernel's array, closed-form outputs (lattice ratios) are expanded here from the plan.
int b200dsp_bank_fetch_schedule(b200dsp_bank_t* b, int chan_id, int32_t* idx, int32_t* phase, int64_t cap, int64_t* n)
{
    if (!b || chan_id < 0 || chan_id >= (int) b->chans.size()) return b200_fail(B200DSP_EINVAL, "bank_fetch_schedule: bad channel");
    Channel& c = b->chans[chan_id];
    if (!c.fe) return b200_fail(B200DSP_ESTATE, "bank_fetch_schedule: channel has no front-end");
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(b->stream)))) return rc;
    if (n) *n = 0;
    if (!b->built || !c.d_plan || !c.d_sched) return 0;
    long long plan[4];
    if ((rc = B200_CUDA_CHECK(cudaMemcpy(plan, c.d_plan, sizeof(plan), cudaMemcpyDeviceToHost)))) return rc;
    const long long k0 = plan[0], i0 = plan[1], D0 = plan[2], ncf = plan[3], total = k0 + ncf;
    if (n) *n = total;
    if (total > cap) return b200_fail(B200DSP_EINVAL, "bank_fetch_schedule: buffer too small (%lld needed)", total);
    if (total == 0 || !idx || !phase) return 0;
    std::vector<int> raw((size_t) k0);
    if (k0 > 0 && (rc = B200_CUDA_CHECK(cudaMemcpy(raw.data(), c.d_sched, (size_t) k0 * sizeof(int), cudaMemcpyDeviceToHost)))) return rc;
    for (long long o = 0; o < k0; ++o) { idx[o] = (int32_t) ((unsigned) raw[(size_t) o] >> 8); phase[o] = raw[(size_t) o] & 0xff; }
    for (long long o = k0; o < total; ++o) {
        const long long E = D0 + (o - k0) * c.A;                  // units of 2^-23 inputs (frontend.cuh: closed-form region)
        idx[o] = (int32_t) (i0 + (E >> 23) - 1);
        phase[o] = (int32_t) ((E & 0x7fffffll) >> c.phshift);
    }
    return 0;
}

int b200dsp_bank_fetch_dev(b200dsp_bank_t* b, int chan_id, int stage, const void** d_ptr, int64_t* n)
{
    if (!b || chan_id < 0 || chan_id >= (int) b->chans.size() || !d_ptr) return b200_fail(B200DSP_EINVAL, "bank_fetch_dev: bad argument");
    Channel& c = b->chans[chan_id];
    if (stage == B200DSP_STAGE_CHANNELIZER) { *d_ptr = c.d_out; if (n) *n = c.out_count; return 0; }
    if (stage == B200DSP_STAGE_FRONTEND && c.fe) { *d_ptr = c.d_fe_out; if (n) *n = -1; return 0; }
    return b200_fail(B200DSP_EINVAL, "bank_fetch_dev: bad stage");
}

// Every channel's outputs of the last feed gathered into one [channel][stride] device array (+ per-channel counts), on the
// stream: the device half of b200dsp_bank_fetch_all, for callers that pipeline the device-to-host copy themselves.
int b200dsp_bank_gather_dev(b200dsp_bank_t* b, int stage, void* d_out, int64_t stride_samples, int64_t* d_counts, void* cuda_stream)
{
    if (!b || !d_out || !d_counts || stride_samples <= 0) return b200_fail(B200DSP_EINVAL, "bank_gather_dev: bad argument");
    if (stage != B200DSP_STAGE_CHANNELIZER && stage != B200DSP_STAGE_FRONTEND) return b200_fail(B200DSP_EINVAL, "bank_gather_dev: bad stage");
    const size_t nc = b->chans.size();
    if (nc == 0) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : b->stream;
    if (b->gcap < nc) {
        if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(st)))) return rc;
        if (b->d_gsrc) cudaFree(b->d_gsrc);
        if (b->d_gcnt) cudaFree(b->d_gcnt);
        b->d_gsrc = nullptr; b->d_gcnt = nullptr; b->gcap = 0;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&b->d_gsrc, 2 * nc * sizeof(GatherSrc)))) || (rc = B200_CUDA_CHECK(cudaMalloc(&b->d_gcnt, nc * sizeof(long long))))) return rc;
        b->gcap = nc;
    }
    // the source table changes only with a reallocation or a new feed's channelizer counts: two slots, alternating, so a
    // table still being read by an earlier gather on another stream is not overwritten
    b->gslot ^= 1;
    GatherSrc* d_tab = b->d_gsrc + (size_t) b->gslot * nc;
    b->h_gsrc.resize(nc);
    long long max_count = 0;
    for (size_t i = 0; i < nc; ++i) {
        const Channel& c = b->chans[i];
        GatherSrc& g = b->h_gsrc[i];
        if (stage == B200DSP_STAGE_CHANNELIZER) { g.src = c.d_out; g.state = nullptr; g.count = b->built ? c.out_count : 0; }
        else if (c.fe && b->built && c.d_fe_out) { g.src = c.d_fe_out; g.state = c.d_state; g.count = 0; }
        else { g.src = nullptr; g.state = nullptr; g.count = 0; }
        if (c.out_count > max_count) max_count = c.out_count;       // front-end outputs never exceed the channel's input count when decimating
    }
    if (max_count > stride_samples) max_count = stride_samples;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(d_tab, b->h_gsrc.data(), nc * sizeof(GatherSrc), cudaMemcpyHostToDevice, st)))) return rc;
    int gx = (int) ((max_count + 1023) / 1024);
    if (gx < 1) gx = 1;
    if (gx > 256) gx = 256;
    if (stage == B200DSP_STAGE_FRONTEND) gather_outputs_kernel<float2><<<dim3(gx, (unsigned) nc), 256, 0, st>>>(d_tab, (float2*) d_out, stride_samples, (long long*) d_counts);
    else gather_outputs_kernel<uint32_t><<<dim3(gx, (unsigned) nc), 256, 0, st>>>(d_tab, (uint32_t*) d_out, stride_samples, (long long*) d_counts);
    return B200_CUDA_CHECK(cudaGetLastError());
}

namespace {

// per-channel output counts of the last feed for `stage` (front-end counts live on the device: one small copy)
int stage_counts(b200dsp_bank* b, int stage, int64_t* counts, cudaStream_t st)
{
    const size_t nc = b->chans.size();
    for (size_t i = 0; i < nc; ++i) counts[i] = 0;
    if (!b->built) return 0;
    if (stage == B200DSP_STAGE_CHANNELIZER) { for (size_t i = 0; i < nc; ++i) counts[i] = b->chans[i].out_count; return 0; }
    const size_t nfe = b->fe_index.size();
    if (!nfe || !b->slab_fe) return 0;
    std::vector<int> stt(nfe * 4);
    int rc;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(stt.data(), b->slab_state, nfe * 4 * sizeof(int), cudaMemcpyDeviceToHost, st))) ||
        (rc = B200_CUDA_CHECK(cudaStreamSynchronize(st)))) return rc;
    // (a channel's counters are its schedule leader's: c.d_state points into the leader's slot)
    for (size_t k = 0; k < nfe; ++k) counts[b->fe_index[k]] = stt[(size_t) (b->chans[b->fe_index[k]].d_state - b->slab_state) + 3];
    return 0;
}

// columns [c0, c1) of every channel's output row -> out (row pitch stride_samples), asynchronously
int copy_columns(b200dsp_bank* b, int stage, void* out, int64_t stride_samples, long long c0, long long c1, cudaStream_t st)
{
    if (c1 <= c0) return 0;
    const size_t elem = (stage == B200DSP_STAGE_FRONTEND) ? sizeof(float2) : sizeof(uint32_t);
    const char* src = (stage == B200DSP_STAGE_FRONTEND) ? (const char*) b->slab_fe : (const char*) b->slab_out;
    const long long pitch = (stage == B200DSP_STAGE_FRONTEND) ? b->fe_pitch : b->out_pitch;
    if (!src) return 0;
    return B200_CUDA_CHECK(cudaMemcpy2DAsync((char*) out + (size_t) c0 * elem, (size_t) stride_samples * elem, src + (size_t) c0 * elem, (size_t) pitch * elem,
                                             (size_t) (c1 - c0) * elem, b->chans.size(), cudaMemcpyDeviceToHost, st));
}

} // namespace

// Every channel's outputs of the last feed in one strided device-to-host transfer (the per-channel b200dsp_bank_fetch costs
// a synchronous small copy per channel: latency-bound for a 1024-channel bank).  The channels' output buffers are rows of
// one device slab, so this is a single 2-D copy of the occupied columns.
int b200dsp_bank_fetch_all(b200dsp_bank_t* b, int stage, void* out, int64_t stride_samples, int64_t* counts, void* cuda_stream)
{
    if (!b || !out || !counts || stride_samples <= 0) return b200_fail(B200DSP_EINVAL, "bank_fetch_all: bad argument");
    if (stage != B200DSP_STAGE_CHANNELIZER && stage != B200DSP_STAGE_FRONTEND) return b200_fail(B200DSP_EINVAL, "bank_fetch_all: bad stage");
    const size_t nc = b->chans.size();
    if (nc == 0) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : b->stream;
    if ((rc = stage_counts(b, stage, counts, st))) return rc;
    long long mx = 0;
    for (size_t i = 0; i < nc; ++i) {
        if (counts[i] > stride_samples) return b200_fail(B200DSP_EINVAL, "bank_fetch_all: stride too small (channel %d has %lld samples)", (int) i, (long long) counts[i]);
        if (counts[i] > mx) mx = counts[i];
    }
    if ((rc = copy_columns(b, stage, out, stride_samples, 0, mx, st))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(st));
}

// == one engine work cycle for the whole bank, host to host (what DSPDeviceSourceEngine::work does with one FIFO read:
// dspdevicesourceengine.cpp:325-408): feed n samples from a host buffer and receive every channel's outputs of `stage` at
// out + c * stride_samples, counts[c] samples each -- the same results as b200dsp_bank_feed followed by b200dsp_bank_fetch_all.
// The block runs as passes of `chunk` samples over three streams: the host-to-device copy of pass p+1, the kernels of pass
// p, and the device-to-host copy of the output columns every channel has completed so far (a strided 2-D copy straight
// out of the channels' slab) overlap; the ragged last columns follow after the last pass.  Pin `iq` and `out`
// (cudaHostRegister) for the copies to be asynchronous.
int b200dsp_bank_process(b200dsp_bank_t* b, const int16_t* iq, int64_t n_samples, int stage, void* out, int64_t stride_samples, int64_t* counts)
{
    if (!b || !out || !counts || stride_samples <= 0) return b200_fail(B200DSP_EINVAL, "bank_process: bad argument");
    if (stage != B200DSP_STAGE_CHANNELIZER && stage != B200DSP_STAGE_FRONTEND) return b200_fail(B200DSP_EINVAL, "bank_process: bad stage");
    if (n_samples < 0 || (n_samples > 0 && !iq)) return b200_fail(B200DSP_EINVAL, "bank_process: bad buffer");
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    if (!b->built && (rc = build(b))) return rc;
    if ((rc = reserve_outputs(b, n_samples))) return rc;
    const size_t nc = b->chans.size();
    for (auto& c : b->chans) c.out_count = 0;
    b->out_count_depth.assign(32, 0);
    if (n_samples == 0) { for (size_t i = 0; i < nc; ++i) counts[i] = 0; return empty_feed(b, b->stream); }
    if ((rc = ensure_root_buffers(b, n_samples))) return rc;
    long long done = 0, copied = 0;
    int pass = 0;
    while (done < n_samples) {
        const long long m = next_pass_len(b, n_samples - done);
        const int pend = (b->depth >= 1) ? (int) (b->produced[0] - 2 * b->produced[1]) : 0;
        const int slot = pass & 1;
        if (pass >= 2 && (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(b->copy, b->ev_done[slot], 0)))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(b->d_root + pend, iq + 2 * done, (size_t) m * 4, cudaMemcpyHostToDevice, b->copy))) ||
            (rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_h2d[slot], b->copy))) ||
            (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(b->stream, b->ev_h2d[slot], 0)))) return rc;
        if ((rc = feed_chunk(b, b->d_root + pend, m, b->stream, done == 0))) return rc;
        if ((rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_done[slot], b->stream)))) return rc;
        std::swap(b->d_root, b->d_root_alt);
        std::swap(b->root_cap, b->root_alt_cap);
        done += m;
        ++pass;
        // columns every channel is certain to hold by now (front-end: a safe lower bound from its input count and ratio)
        long long ready = -1;
        for (size_t i = 0; i < nc; ++i) {
            const Channel& c = b->chans[i];
            long long r;
            if (stage == B200DSP_STAGE_CHANNELIZER) r = b->out_count_depth[c.S];
            else if (!c.fe) continue;
            else r = (long long) ((double) b->out_count_depth[c.S] / (double) c.ratio) - 2;
            if (ready < 0 || r < ready) ready = r;
        }
        if (ready > stride_samples) ready = stride_samples;
        if (ready > copied && done < n_samples) {
            if ((rc = B200_CUDA_CHECK(cudaEventRecord(b->ev_pass, b->stream))) || (rc = B200_CUDA_CHECK(cudaStreamWaitEvent(b->d2h, b->ev_pass, 0))) ||
                (rc = copy_columns(b, stage, out, stride_samples, copied, ready, b->d2h))) return rc;
            copied = ready;
        }
    }
    if ((rc = B200_CUDA_CHECK(cudaStreamSynchronize(b->stream)))) return rc;
    if ((rc = stage_counts(b, stage, counts, b->stream))) return rc;
    long long mx = 0;
    for (size_t i = 0; i < nc; ++i) {
        if (counts[i] > stride_samples) { cudaStreamSynchronize(b->d2h); return b200_fail(B200DSP_EINVAL, "bank_process: stride too small (channel %d has %lld samples)", (int) i, (long long) counts[i]); }
        if (counts[i] > mx) mx = counts[i];
    }
    if ((rc = copy_columns(b, stage, out, stride_samples, copied, mx, b->d2h))) return rc;
    return B200_CUDA_CHECK(cudaStreamSynchronize(b->d2h));
}

void* b200dsp_bank_stream(b200dsp_bank_t* b) { return b ? (void*) b->stream : nullptr; }

// device-to-device copy of part of a channel's output of the last feed (asynchronous on the stream): the building block of
// the cooperative multi-GPU bank, where one bank's node outputs are exchanged and become another bank's input
int b200dsp_bank_copy_out_dev(b200dsp_bank_t* b, int chan_id, int64_t skip, int64_t count, void* d_dst, void* cuda_stream)
{
    if (!b || chan_id < 0 || chan_id >= (int) b->chans.size() || skip < 0 || count < 0 || (count > 0 && !d_dst)) return b200_fail(B200DSP_EINVAL, "bank_copy_out_dev: bad argument");
    Channel& c = b->chans[chan_id];
    if (skip + count > c.out_count) return b200_fail(B200DSP_EINVAL, "bank_copy_out_dev: range beyond the %lld samples of the last feed", c.out_count);
    if (count == 0) return 0;
    int rc = B200_CUDA_CHECK(cudaSetDevice(b->device));
    if (rc) return rc;
    return B200_CUDA_CHECK(cudaMemcpyAsync(d_dst, c.d_out + skip, (size_t) count * 4, cudaMemcpyDeviceToDevice, cuda_stream ? (cudaStream_t) cuda_stream : b->stream));
}

int b200dsp_bank_sync(b200dsp_bank_t* b)
{
    if (!b) return b200_fail(B200DSP_EINVAL, "null handle");
    return B200_CUDA_CHECK(cudaStreamSynchronize(b->stream));
}


// ---------------------------------------------------------------------------------------------------------
// Stand-alone Interpolator (sdrbase/dsp/interpolator.h:19-36, interpolator.cpp:74-129): the block form of the loop every
// Rx plugin writes around Interpolator::decimate (nfmdemod.cpp:150-155,315).  Same kernels as the bank's front-end,
// complex64 input, no NCO.
// ---------------------------------------------------------------------------------------------------------
int b200dsp_interp_create(b200dsp_interp_t** out, int phase_steps, double sample_rate, double cutoff, double taps_per_phase)
{
    if (!out) return b200_fail(B200DSP_EINVAL, "interp_create: null handle pointer");
    *out = nullptr;
    if (phase_steps < 1 || phase_steps > 255 || sample_rate <= 0 || taps_per_phase <= 0) return b200_fail(B200DSP_EINVAL, "interp_create: bad parameters");
    int rc = b200_require_device();
    if (rc) return rc;
    b200dsp_interp* h = new (std::nothrow) b200dsp_interp();
    if (!h) return b200_fail(B200DSP_ENOMEM, "interp_create: out of host memory");
    h->device = b200_current_device();
    h->phase_steps = phase_steps;
    interp_taps(phase_steps, sample_rate, cutoff, taps_per_phase, h->taps, &h->ntaps);
    if (h->ntaps > FE_MAX_TAPS) { delete h; return b200_fail(B200DSP_EINVAL, "interp_create: %d taps per phase exceed the supported %d", h->ntaps, FE_MAX_TAPS); }
    const size_t hist_bytes = 2 * (2 * FE_MAX_TAPS + 4) * 4;
    if ((rc = B200_CUDA_CHECK(cudaSetDevice(h->device))) || (rc = B200_CUDA_CHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_taps, h->taps.size() * 4))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpy(h->d_taps, h->taps.data(), h->taps.size() * 4, cudaMemcpyHostToDevice))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_hist, hist_bytes))) || (rc = B200_CUDA_CHECK(cudaMemset(h->d_hist, 0, hist_bytes))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_state, 16))) || (rc = B200_CUDA_CHECK(cudaMemset(h->d_state, 0, 16))) ||
        (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_plan, 32))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_chan, sizeof(FrontendChan)))) ||
        (rc = B200_CUDA_CHECK(cudaDeviceSynchronize()))) { b200dsp_interp_destroy(h); return rc; }
    *out = h;
    return 0;
}

int b200dsp_interp_destroy(b200dsp_interp_t* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* ptrs[] = { h->d_taps, h->d_hist, h->d_state, h->d_plan, h->d_chan, h->d_in, h->d_out, h->d_sched, h->d_tile };
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int b200dsp_interp_info(b200dsp_interp_t* h, int* taps_per_phase, float* taps, int taps_cap)
{
    if (!h) return b200_fail(B200DSP_EINVAL, "null handle");
    if (taps_per_phase) *taps_per_phase = h->ntaps;
    if (taps) for (int i = 0; i < (int) h->taps.size() && i < taps_cap; ++i) taps[i] = h->taps[i];
    return 0;
}

namespace {

// the three caller loops around Interpolator (frontend.cuh: FrontendChan::mode) on one block of complex64 samples
int interp_run(b200dsp_interp* h, int mode, float* distance_remain, float distance, const float* in_c64, int64_t n, float* out_c64, int64_t cap, int64_t* n_out,
               bool single = false)
{
    if (!h || !distance_remain) return b200_fail(B200DSP_EINVAL, "interp: null argument");
    if (n < 0 || n >= (1 << 24) - 1 || (n > 0 && !in_c64) || (cap > 0 && !out_c64) || cap < 0) return b200_fail(B200DSP_EINVAL, "interp: bad buffer (at most 2^24-2 samples per call)");
    if (!(distance > 0.0f) && !single) return b200_fail(B200DSP_EINVAL, "interp: distance must be positive");
    if (n_out) *n_out = 0;
    if (n == 0 && mode != 1) return 0;                       // (the interpolate loop may still emit outputs while distance_remain < 1)
    int rc = B200_CUDA_CHECK(cudaSetDevice(h->device));
    if (rc) return rc;
    // outputs of one call: never more than the inputs when decimating; otherwise bounded by (n + 2) / distance + 2
    long long cap_out = n + 2;
    if (mode != 0 && !single) {
        const double bound = ((double) n + 2.0) / (double) distance + 4.0;
        if (bound >= (double) (1 << 28)) return b200_fail(B200DSP_EINVAL, "interp: distance too small for one call");
        if ((long long) bound > cap_out) cap_out = (long long) bound;
    }
    if (h->cap < n || h->cap_out < cap_out) {
        void* old[] = { h->d_in, h->d_out, h->d_sched, h->d_tile };
        for (void* p : old) if (p) cudaFree(p);
        h->d_in = nullptr; h->d_out = nullptr; h->d_sched = nullptr; h->d_tile = nullptr; h->cap = 0; h->cap_out = 0;
        const long long ci = n > h->cap ? n : h->cap, co = cap_out > h->cap_out ? cap_out : h->cap_out;
        if ((rc = B200_CUDA_CHECK(cudaMalloc(&h->d_in, (size_t) (ci + 1) * 8))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_out, (size_t) (co + 2) * 8))) ||
            (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_sched, (size_t) (co + 2) * 4))) || (rc = B200_CUDA_CHECK(cudaMalloc(&h->d_tile, (size_t) (ci / FE_TILE + 4) * 4)))) return rc;
        h->cap = ci; h->cap_out = co;
    }
    FrontendChan f;
    memset(&f, 0, sizeof(f));
    f.in = (const uint32_t*) h->d_in; f.hist = h->d_hist; f.taps = h->d_taps; f.out = h->d_out; f.sched = h->d_sched; f.tile_start = h->d_tile;
    f.state = h->d_state; f.plan = h->d_plan; f.in_f32 = 1; f.hist_stride = 2 * FE_MAX_TAPS + 4;
    f.depth = 0; f.inc = 0; f.ntaps = h->ntaps; f.phase_steps = h->phase_steps; f.ratio = distance;
    f.mode = mode; f.sched_cap = single ? 1 : (int) h->cap_out;
    f.lattice = (mode == 0 && lattice_params(distance, h->phase_steps, &f.A, &f.phshift)) ? 1 : 0;
    if (mode == 0 && !f.lattice && !single && !getenv("B200DSP_SERIAL_SCHEDULE")) { long long a = 0; int kb = 0; if (scan_params(distance, &kb, &a)) { f.scan_kb = kb; f.A = a; } }
    // the caller owns the distance (Real* distance in the reference): it travels in, and back out
    int st[4] = { 0, 0, 0, 0 };
    memcpy(&st[1], distance_remain, 4);
    // an off-lattice remain (set by the caller) makes the closed form invalid: the serial replay is always exact
    if (f.lattice && ((double) *distance_remain * 8388608.0 != floor((double) *distance_remain * 8388608.0) ||
                      ((long long) ((double) *distance_remain * 8388608.0)) % (f.A & -f.A ? (f.A & -f.A) : 1) != 0)) f.lattice = 0;
    PassInfo pi;
    memset(&pi, 0, sizeof(pi));
    pi.n_new[0] = (int) n; pi.first_pass = 1; pi.parity = h->parity;
    const size_t smem = ((((size_t) ((h->ntaps + 2 * FE_PAD) | 1) * h->phase_steps + 3) & ~(size_t) 3)) * 4 + (size_t) (FE_MAX_TAPS + FE_TILE + FE_Z_EXTRA) * sizeof(float2);
    if ((n > 0 && (rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_in, in_c64, (size_t) n * 8, cudaMemcpyHostToDevice, h->stream)))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_state, st, 16, cudaMemcpyHostToDevice, h->stream))) ||
        (rc = B200_CUDA_CHECK(cudaMemcpyAsync(h->d_chan, &f, sizeof(f), cudaMemcpyHostToDevice, h->stream)))) return rc;
    if (smem > 48 * 1024 && (rc = B200_CUDA_CHECK(cudaFuncSetAttribute((const void*) frontend_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)))) return rc;
    frontend_schedule_kernel<<<1, 32 * FE_SW, 0, h->stream>>>(h->d_chan, 1, pi);
    const unsigned tiles = (unsigned) ((n + FE_TILE - 1) / FE_TILE);
    frontend_kernel_t<false><<<dim3(tiles ? tiles : 1, 1), FE_THREADS, smem, h->stream>>>(h->d_chan, nullptr, pi);
    if ((rc = B200_CUDA_CHECK(cudaGetLastError()))) return rc;
    if ((rc = B200_CUDA_CHECK(cudaMemcpyAsync(st, h->d_state, 16, cudaMemcpyDeviceToHost, h->stream))) ||
        (rc = B200_CUDA_CHECK(cudaStreamSynchronize(h->stream)))) return rc;
    h->parity ^= 1;
    if (st[0] && !single) return b200_fail(B200DSP_ESTATE, "interp: more outputs than the schedule buffer holds (internal bound exceeded)");
    memcpy(distance_remain, &st[1], 4);
    const long long m = st[2];
    if (n_out) *n_out = m;
    if (m > cap) return b200_fail(B200DSP_EINVAL, "interp: output buffer too small (%lld needed)", m);
    if (m > 0) return B200_CUDA_CHECK(cudaMemcpy(out_c64, h->d_out, (size_t) m * 8, cudaMemcpyDeviceToHost));
    return 0;
}

} // namespace

int b200dsp_interp_decimate(b200dsp_interp_t* h, float* distance_remain, float distance, const float* in_c64, int64_t n, float* out_c64, int64_t cap, int64_t* n_out)
{
    return interp_run(h, 0, distance_remain, distance, in_c64, n, out_c64, cap, n_out);
}

int b200dsp_interp_interpolate(b200dsp_interp_t* h, float* distance_remain, float distance, const float* in_c64, int64_t n, float* out_c64, int64_t cap, int64_t* n_out)
{
    return interp_run(h, 1, distance_remain, distance, in_c64, n, out_c64, cap, n_out);
}

int b200dsp_interp_resample(b200dsp_interp_t* h, float* distance_remain, float distance, const float* in_c64, int64_t n, float* out_c64, int64_t cap, int64_t* n_out)
{
    return interp_run(h, 2, distance_remain, distance, in_c64, n, out_c64, cap, n_out);
}

// One reference call, for plugin code that has not been restructured into blocks (one launch per sample: drop-in, not fast):
//   op 0: Interpolator::decimate(distance, next, result)            -> *produced = its return value (a result was written)
//   op 1: Interpolator::interpolate(distance, next, result)         -> *produced = 1, *consumed = its return value
//   op 2: Interpolator::resample(distance, next, consumed, result)  -> *produced = its return value, *consumed in/out
// `distance` is updated exactly as the reference method does (the caller adds its step afterwards, as in the reference).
int b200dsp_interp_step(b200dsp_interp_t* h, int op, float* distance, const float* next_c64, float* result_c64, int* consumed, int* produced)
{
    if (!h || !distance || !next_c64 || !result_c64 || !produced || op < 0 || op > 2 || (op != 0 && !consumed)) return b200_fail(B200DSP_EINVAL, "interp_step: bad argument");
    int64_t m = 0;
    int rc;
    *produced = 0;
    if (op == 0) {
        // push, distance -= 1, a result iff distance < 1 (interpolator.h:23-36): the block form with one input and step 0
        if ((rc = interp_run(h, 0, distance, 0.0f, next_c64, 1, result_c64, 1, &m, true))) return rc;
        *produced = (int) m;
        return 0;
    }
    if (op == 1) {
        // consume iff distance >= 1, then always a result (interpolator.h:39-52)
        const int take = (*distance >= 1.0f) ? 1 : 0;
        if ((rc = interp_run(h, 1, distance, 0.0f, next_c64, take, result_c64, 1, &m, true))) return rc;
        *consumed = take; *produced = 1;
        return 0;
    }
    // resample (interpolator.h:55-76)
    if (*distance >= 1.0f) {
        if (*consumed) return 0;                                   // "return false": nothing happens
        *consumed = 1;
        if (*distance >= 2.0f) return interp_run(h, 0, distance, 0.0f, next_c64, 1, result_c64, 1, &m, true);      // pushed, still >= 1: no result
        if ((rc = interp_run(h, 1, distance, 0.0f, next_c64, 1, result_c64, 1, &m, true))) return rc;
        *produced = 1;
        return 0;
    }
    if ((rc = interp_run(h, 1, distance, 0.0f, next_c64, 0, result_c64, 1, &m, true))) return rc;
    *produced = 1;
    return 0;
}

} // extern "C"
