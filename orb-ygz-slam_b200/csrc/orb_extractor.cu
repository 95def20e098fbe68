// Host side of the batched ORB extractor: geometry / table construction that mirrors the
// ORBextractor constructor and cell loop (src/ORBextractor.cc:412-470, 725-762, 1129-1133), device
// buffers, kernel launches and the extern "C" entry points declared in include/ygz_b200.h.
#include "orb_kernels.cuh"
#include "fast_cells_v2.cuh"
#include "pyr_resize_v2.cuh"
#include "dso_kernels.cuh"
#include "trig_glibc.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace ygz_b200 {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }

static const int8_t h_pattern[256 * 4] = {
#include "orb_pattern.inc"
};

struct LevelHost {
    int w, h, pitch;
    size_t frame_stride;
    uint8_t* d_base = nullptr;
    uint8_t* d_blur = nullptr;     // GaussianBlur(level), written by k_blur_levels
    ResizeTabX* d_tx = nullptr;
    ResizeTabY* d_ty = nullptr;
    int py2_pitch = 0, py2_smem = 0;   // k_pyr_resize_v2 staging window of this level (0: use k_pyr_resize)
};

}  // namespace ygz_b200

using namespace ygz_b200;

struct ygz_orb {
    ygz_orb_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    OrbParams P;
    std::vector<LevelHost> lv;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> feat_per_level;
    std::vector<CellDesc> cells;
    CellDesc* d_cells = nullptr;
    uint32_t* d_pool = nullptr;
    uint16_t* d_cell_count = nullptr;
    uint32_t* d_keys = nullptr;
    uint16_t* d_keynode = nullptr;
    int* d_nkeys = nullptr;
    uint32_t* d_sel = nullptr;
    int* d_nsel = nullptr;
    ygz_keypoint* d_kps = nullptr;
    uint8_t* d_desc = nullptr;
    int* d_counts = nullptr;
    ExistingKey* d_existing = nullptr;
    uint8_t* d_stage = nullptr;          // upload staging for frames whose rows are not 16-byte multiples (see ygz_orb_upload)
    size_t stage_bytes = 0;
    int existing_cap = 0;
    float4* d_pattern = nullptr;
    TmaMaps tma;
    PatchMaps patch;
    FastMaps fastmaps;
    int* d_fast_ids = nullptr;          // [narrow cells (k_fast_cells_v2)] ++ [the rest (k_fast_cells)]
    int* d_fast_queue = nullptr;        // work counter of k_fast_cells_v2 (zeroed before every launch)
    cudaStream_t aux_stream = nullptr;  // the generic FAST kernel (few, wide cells) runs beside k_fast_cells_v2
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int n_fast_narrow = 0, n_fast_wide = 0;
    int fast_v2_grid = 0;
    BlurTile* d_blur_tiles = nullptr;
    int n_blur_tiles = 0;
    size_t fast_smem = 0, oct_smem = 0;
    int fast_warp_smem = 0;
    int oct_node_smem = 0, oct_key_cache = 0, oct_unroll = 4;     // measured (C1, B=256): 4 keys in flight 0.275 ms, 8 keys 0.296 ms (spills), one 0.300 ms
    int launches_per_run = 0;
    ygz_counters ctr = {};
    // DSO key-point selection (ygz_orb_extract_frame_dso): occupancy map, per-cell winners, compacted list, mnGridSize
    uint8_t* d_occ = nullptr;
    uint32_t* d_cell_keys = nullptr;
    uint8_t* d_cell_cnt = nullptr;
    uint32_t* d_dso_keys = nullptr;
    int* d_dso_total = nullptr;
    float2* d_dso_pts = nullptr;
    int dso_pts_cap = 0;
    int dso_grid = -1;
    int dso_frames = 0;                 // frames the DSO scratch (d_occ, d_cell_*, d_dso_keys) is allocated for
    DsoFrameState* d_dso_state = nullptr;
    int* d_dso_fid = nullptr;           // frame id per existing key point of a batch (k_dso_mark_occupancy_batch)
    int dso_fid_cap = 0;
    int last_n = 0;
};

static void build_resize_tables(int ssize, int dsize, std::vector<ResizeTabX>* tx, std::vector<ResizeTabY>* ty) {
    // cv::resize INTER_LINEAR table construction (imgproc/resize.cpp): fx = (float)((dx+0.5)*scale - 0.5),
    // sx = cvFloor(fx), coefficients saturate_cast<short>(w * 2048)
    const double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = host_cv_floor(f);
        f -= s;
        if (tx) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
            ResizeTabX t;
            t.ofs = s;
            t.a0 = (short)host_cv_round((1.f - f) * 2048);
            t.a1 = (short)host_cv_round(f * 2048);
            tx->push_back(t);
        } else {
            ResizeTabY t;   // vertical: fractional weights kept, row indices clipped
            t.b0 = (short)host_cv_round((1.f - f) * 2048);
            t.b1 = (short)host_cv_round(f * 2048);
            t.y0 = (short)std::min(std::max(s, 0), ssize - 1);
            t.y1 = (short)std::min(std::max(s + 1, 0), ssize - 1);
            ty->push_back(t);
        }
    }
}

static int orb_free(ygz_orb* h) {
    if (!h) return YGZ_OK;
    cudaSetDevice(h->device);
    for (auto& L : h->lv) { cudaFree(L.d_base); cudaFree(L.d_blur); cudaFree(L.d_tx); cudaFree(L.d_ty); }
    cudaFree(h->d_blur_tiles); cudaFree(h->d_fast_ids); cudaFree(h->d_fast_queue);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    cudaFree(h->d_occ); cudaFree(h->d_cell_keys); cudaFree(h->d_cell_cnt); cudaFree(h->d_dso_keys); cudaFree(h->d_dso_total); cudaFree(h->d_dso_pts); cudaFree(h->d_dso_state); cudaFree(h->d_dso_fid);
    cudaFree(h->d_cells); cudaFree(h->d_pool); cudaFree(h->d_cell_count); cudaFree(h->d_keys);
    cudaFree(h->d_keynode); cudaFree(h->d_nkeys); cudaFree(h->d_sel); cudaFree(h->d_nsel);
    cudaFree(h->d_kps); cudaFree(h->d_desc); cudaFree(h->d_counts); cudaFree(h->d_existing); cudaFree(h->d_stage); cudaFree(h->d_pattern);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return YGZ_OK;
}

extern "C" {

const char* ygz_last_error(void) { return g_last_error.c_str(); }

int ygz_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ygz_device_info(int dev, char* name, int* sm_count, size_t* l2_bytes) {
    cudaDeviceProp p;
    YGZ_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    if (name) { std::strncpy(name, p.name, 255); name[255] = 0; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
    return YGZ_OK;
}

int ygz_orb_create(const ygz_orb_config* cfg, ygz_orb** out) {
    YGZ_ARG_CHECK(cfg && out);
    YGZ_ARG_CHECK(cfg->nlevels >= 1 && cfg->nlevels <= MAX_LEVELS);
    YGZ_ARG_CHECK(cfg->width >= 64 && cfg->height >= 64 && cfg->width <= 4096 && cfg->height <= 4096);
    YGZ_ARG_CHECK(cfg->max_batch >= 1 && cfg->nfeatures >= 1 && cfg->scale_factor > 1.0f);
    // any order of the two thresholds is legal in the reference (FAST(iniTh), then FAST(minTh) for empty cells)
    YGZ_ARG_CHECK(cfg->ini_th_fast >= 1 && cfg->ini_th_fast <= 255 && cfg->min_th_fast >= 1 && cfg->min_th_fast <= 255);
    if (ygz_device_count() <= cfg->device) { set_last_error("no such CUDA device"); return YGZ_ERR_NO_DEVICE; }
    YGZ_CUDA_TRY(cudaSetDevice(cfg->device));
    ygz_orb* h = new ygz_orb();
    h->cfg = *cfg;
    h->device = cfg->device;
    const int nl = cfg->nlevels, B = cfg->max_batch;
    std::memset(&h->P, 0, sizeof(h->P));

    // ---- scale tables and features per level (src/ORBextractor.cc:416-445); scaleFactor is a double member
    const double sf = (double)cfg->scale_factor;
    h->scale.resize(nl); h->inv_scale.resize(nl); h->sigma2.resize(nl); h->inv_sigma2.resize(nl);
    h->scale[0] = 1.0f; h->sigma2[0] = 1.0f;
    for (int i = 1; i < nl; i++) {
        h->scale[i] = (float)(h->scale[i - 1] * sf);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < nl; i++) { h->inv_scale[i] = 1.0f / h->scale[i]; h->inv_sigma2[i] = 1.0f / h->sigma2[i]; }
    h->feat_per_level.resize(nl);
    {
        float factor = (float)(1.0f / sf);
        float nDesired = (float)(cfg->nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl)));
        int sum = 0;
        for (int l = 0; l < nl - 1; l++) {
            h->feat_per_level[l] = host_cv_round(nDesired);
            sum += h->feat_per_level[l];
            nDesired *= factor;
        }
        h->feat_per_level[nl - 1] = std::max(cfg->nfeatures - sum, 0);
    }
    // ---- umax (src/ORBextractor.cc:453-467)
    {
        int umax[HALF_PATCH + 2] = {0};
        int v, v0, vmax = host_cv_floor((float)(HALF_PATCH * std::sqrt(2.f) / 2 + 1));
        int vmin = (int)std::ceil(HALF_PATCH * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH * HALF_PATCH;
        for (v = 0; v <= vmax; ++v) umax[v] = host_cv_round(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
        for (int i = 0; i < 16; i++) h->P.umax[i] = umax[i];
        for (int u = 0; u < 16; u++) {           // rows covered by column u (umax is non-increasing in v)
            int vl = -1;
            for (int vv = 0; vv <= HALF_PATCH; vv++) if (u <= umax[vv]) vl = vv;
            h->P.vlim[u] = vl;
        }
    }

    // ---- level geometry (:1131-1132), FAST cells (:739-762), octree roots (:537-551)
    h->lv.resize(nl);
    h->P.nlevels = nl;
    h->P.ini_th = cfg->ini_th_fast;
    h->P.min_th = cfg->min_th_fast;
    int pool = 0, key_total = 0, sel_total = 0;
    size_t fast_smem = 0, oct_smem = 0;
    for (int l = 0; l < nl; l++) {
        LevelHost& L = h->lv[l];
        L.w = host_cv_round((float)cfg->width * h->inv_scale[l]);
        L.h = host_cv_round((float)cfg->height * h->inv_scale[l]);
        if (L.w < 1 || L.h < 1) { orb_free(h); set_last_error("pyramid level vanishes"); return YGZ_ERR_ARG; }
        L.pitch = round_up(L.w, 16);     // 16-byte vector access; 752 stays 752 so contiguous host frames upload as one 1-D copy
        L.frame_stride = (size_t)L.pitch * L.h;
        LevelDev& D = h->P.lv[l];
        D.w = L.w; D.h = L.h; D.pitch = L.pitch; D.frame_stride = L.frame_stride;
        D.scale = h->scale[l];
        D.size = (float)(int)(31 * h->scale[l]);
        D.cell_begin = (int)h->cells.size();
        D.key_off = key_total;
        const int minB = MIN_BORDER, maxBX = L.w - EDGE_THRESHOLD + 3, maxBY = L.h - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        const float W = 30;
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        int key_cap = 0;
        if (nCols > 0 && nRows > 0) {       // else: reference divides by zero; level yields nothing
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minB + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minB + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    CellDesc c;
                    c.level = (short)l;
                    c.x0 = (short)iniX; c.y0 = (short)iniY;
                    c.cw = (short)((int)maxX - (int)iniX); c.ch = (short)((int)maxY - (int)iniY);
                    c.offx = (short)(j * wCell); c.offy = (short)(i * hCell);
                    const int dw = std::max(c.cw - 6, 0), dh = std::max(c.ch - 6, 0);
                    c.cap = (short)(((dw + 1) / 2) * ((dh + 1) / 2));
                    c.out_off = pool;
                    pool += c.cap;
                    key_cap += c.cap;
                    h->cells.push_back(c);
                    if (dw > 63 || dh > 63) { orb_free(h); set_last_error("FAST cell larger than 63x63 detection window"); return YGZ_ERR_ARG; }
                }
            }
        }
        D.cell_end = (int)h->cells.size();
        D.key_cap = key_cap;
        key_total += key_cap;
        // DistributeOctTree(minX=16, maxX, minY=16, maxY, N)
        D.n_target = h->feat_per_level[l];
        const int dx = maxBX - minB, dy = maxBY - minB;
        int nIni = (dy > 0 && dx > 0) ? (int)std::round(static_cast<float>(dx) / dy) : 0;
        if (nIni > MAX_ROOTS) { orb_free(h); set_last_error("aspect ratio > 16:1 unsupported"); return YGZ_ERR_ARG; }
        D.n_roots = nIni;
        D.root_h = dy;
        if (nIni >= 1) {
            const float hX = static_cast<float>(dx) / nIni;
            D.hX = hX;
            for (int i = 0; i < nIni; i++) {
                D.root_x0[i] = (short)(int)(hX * static_cast<float>(i));
                D.root_x1[i] = (short)(int)(hX * static_cast<float>(i + 1));
            }
        }
        D.node_cap = std::max(D.n_target + 4, 4 * std::max(nIni, 1) + 4);
        if (D.node_cap > 32000) { orb_free(h); set_last_error("nfeatures per level too large"); return YGZ_ERR_ARG; }
        D.sel_off = sel_total;
        sel_total += D.node_cap;
        oct_smem = std::max(oct_smem, oct_smem_bytes(D.node_cap));
    }
    h->P.n_cells = (int)h->cells.size();
    h->P.pool_size = std::max(pool, 1);
    h->P.key_total = std::max(key_total, 1);
    h->P.sel_total = sel_total;
    h->P.out_cap = sel_total;
    for (const CellDesc& c : h->cells)
        if (c.cw > 6 && c.ch > 6) fast_smem = std::max(fast_smem, (size_t)fast_warp_smem(c.cw, c.ch));
    h->fast_warp_smem = (int)round_up((int)fast_smem, 16);
    fast_smem = (size_t)h->fast_warp_smem * FAST_WARPS;
    h->fast_smem = fast_smem;
    // octree: node tables + a key cache sized for the largest level's realistic candidate count (capacity / 4, <= 12288 keys)
    h->oct_node_smem = (int)round_up((int)oct_smem, 16);
    {
        int kc = 0;
        for (int l = 0; l < nl; l++) kc = std::max(kc, h->P.lv[l].key_cap / 4);
        (void)kc;
        h->oct_key_cache = 0;     // measured on B200 (C1, B=256): 0 keys 0.304 ms, 4608 keys 0.312 ms, 9216 keys 0.338 ms — occupancy beats the cache
        if (const char* e = getenv("YGZ_OCT_KEY_CACHE")) h->oct_key_cache = std::max(0, atoi(e));   // tuning knob
        if (const char* e = getenv("YGZ_OCT_UNROLL")) h->oct_unroll = atoi(e) == 8 ? 8 : 4;           // tuning knob
    }
    oct_smem = (size_t)h->oct_node_smem + (size_t)h->oct_key_cache * 6 + 16;
    h->oct_smem = oct_smem;
    if (fast_smem > 200 * 1024 || oct_smem > 200 * 1024) { orb_free(h); set_last_error("shared memory budget exceeded"); return YGZ_ERR_ARG; }

    // ---- device buffers
#define TRY_OR_FREE(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e)); orb_free(h); return YGZ_ERR_CUDA; } } while (0)
    TRY_OR_FREE(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) TRY_OR_FREE(cudaEventCreate(&e));
    int py2_max_smem = 0;
    for (int l = 0; l < nl; l++) {
        LevelHost& L = h->lv[l];
        TRY_OR_FREE(cudaMalloc(&L.d_base, L.frame_stride * B + 64));
        h->P.lv[l].base = L.d_base;
        TRY_OR_FREE(cudaMalloc(&L.d_blur, L.frame_stride * B + 64));
        h->P.lv[l].blur = L.d_blur;
        if (l > 0) {
            std::vector<ResizeTabX> tx;
            std::vector<ResizeTabY> ty;
            build_resize_tables(h->lv[l - 1].w, L.w, &tx, nullptr);
            build_resize_tables(h->lv[l - 1].h, L.h, nullptr, &ty);
            TRY_OR_FREE(cudaMalloc(&L.d_tx, tx.size() * sizeof(ResizeTabX)));
            TRY_OR_FREE(cudaMalloc(&L.d_ty, ty.size() * sizeof(ResizeTabY)));
            TRY_OR_FREE(cudaMemcpy(L.d_tx, tx.data(), tx.size() * sizeof(ResizeTabX), cudaMemcpyHostToDevice));
            TRY_OR_FREE(cudaMemcpy(L.d_ty, ty.data(), ty.size() * sizeof(ResizeTabY), cudaMemcpyHostToDevice));
            h->P.lv[l].tx = L.d_tx;
            h->P.lv[l].ty = L.d_ty;
            // staging window of k_pyr_resize_v2: the largest source window any 256 x 16 output tile of this level needs
            int max_rows = 0, max_vec = 0;
            for (int y = 0; y < L.h; y += PY2_TILE_ROWS)
                max_rows = std::max(max_rows, ty[std::min(y + PY2_TILE_ROWS, L.h) - 1].y1 - ty[y].y0 + 1);
            for (int x = 0; x < L.w; x += PY2_COLS) {
                const int lo = tx[x].ofs & ~15, hi = std::min(tx[std::min(x + PY2_COLS, L.w) - 1].ofs + 1, h->lv[l - 1].w - 1);
                max_vec = std::max(max_vec, (hi - lo) / 16 + 1);
            }
            const int smem = max_rows * max_vec * 32;
            if (smem <= 96 * 1024 && !getenv("YGZ_PYR_V1")) {
                L.py2_pitch = max_vec * 32;
                L.py2_smem = smem;
                py2_max_smem = std::max(py2_max_smem, smem);
            }
        }
    }
    if (py2_max_smem > 0)
        TRY_OR_FREE(cudaFuncSetAttribute(k_pyr_resize_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(py2_max_smem, 1024)));
    TRY_OR_FREE(cudaMalloc(&h->d_cells, std::max<size_t>(h->cells.size(), 1) * sizeof(CellDesc)));
    if (!h->cells.empty())
        TRY_OR_FREE(cudaMemcpy(h->d_cells, h->cells.data(), h->cells.size() * sizeof(CellDesc), cudaMemcpyHostToDevice));
    TRY_OR_FREE(cudaMalloc(&h->d_pool, (size_t)h->P.pool_size * B * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_cell_count, (size_t)std::max(h->P.n_cells, 1) * B * 2));
    TRY_OR_FREE(cudaMalloc(&h->d_keys, (size_t)h->P.key_total * B * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_keynode, (size_t)h->P.key_total * B * 2));
    TRY_OR_FREE(cudaMalloc(&h->d_nkeys, (size_t)nl * B * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_sel, (size_t)h->P.sel_total * B * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_nsel, (size_t)nl * B * 4));
    TRY_OR_FREE(cudaMalloc(&h->d_kps, (size_t)h->P.out_cap * B * sizeof(ygz_keypoint)));
    TRY_OR_FREE(cudaMalloc(&h->d_desc, (size_t)h->P.out_cap * B * 32));
    TRY_OR_FREE(cudaMalloc(&h->d_counts, (size_t)B * 4));
    TRY_OR_FREE(cudaMemset(h->d_counts, 0, (size_t)B * 4));
    TRY_OR_FREE(cudaMemset(h->d_nsel, 0, (size_t)nl * B * 4));
    // TMA descriptors of the raw pyramid levels for k_blur_levels (box 160 x 38 x 1).  A level narrower / lower than
    // the box, or a failed encode, falls back to the manual loader for that level; YGZ_NO_TMA=1 forces that (A/B runs).
    std::memset(&h->tma, 0, sizeof(h->tma));
    std::memset(&h->patch, 0, sizeof(h->patch));
    {
        std::vector<BlurTile> tiles;
        for (int l = 0; l < nl; l++) {
            const LevelHost& L = h->lv[l];
            if (!getenv("YGZ_NO_TMA") && L.w >= BLUR_BOXW && L.h >= BLUR_BOXH &&
                make_tensor_map_u8_3d(&h->tma.level[l], L.d_base, L.w, L.h, B, L.pitch, L.frame_stride, BLUR_BOXW, BLUR_BOXH) == 0)
                h->tma.ok |= 1u << l;
            if (!getenv("YGZ_NO_TMA") && L.w >= PBOX && L.h >= PW &&
                make_tensor_map_u8_3d(&h->patch.blur[l], L.d_blur, L.w, L.h, B, L.pitch, L.frame_stride, PBOX, PW) == 0 &&
                make_tensor_map_u8_3d(&h->patch.raw[l], L.d_base, L.w, L.h, B, L.pitch, L.frame_stride, PBOX, ICW) == 0)
                h->patch.ok |= 1u << l;
            for (int ty = 0; ty < ceil_div(L.h, BLUR_TH); ty++)
                for (int tx = 0; tx < ceil_div(L.w, BLUR_TW); tx++) tiles.push_back(BlurTile{(short)l, (short)tx, (short)ty, 0});
        }
        h->n_blur_tiles = (int)tiles.size();
        TRY_OR_FREE(cudaMalloc(&h->d_blur_tiles, tiles.size() * sizeof(BlurTile)));
        TRY_OR_FREE(cudaMemcpy(h->d_blur_tiles, tiles.data(), tiles.size() * sizeof(BlurTile), cudaMemcpyHostToDevice));
    }
    {
        // FAST cells: k_fast_cells_v2 (TMA box per cell, one lane per detection column) takes the cells with a detection window
        // of at most 32 columns and a cell image of at most F2_MAX_CH rows on levels that have a tensor map; k_fast_cells the rest.
        std::memset(&h->fastmaps, 0, sizeof(h->fastmaps));
        std::vector<int> narrow, wide;
        unsigned ok = 0;
        for (int l = 0; l < nl; l++) {
            const LevelHost& L = h->lv[l];
            int maxch = 0;
            for (int c = h->P.lv[l].cell_begin; c < h->P.lv[l].cell_end; c++)
                if (h->cells[c].cw - 6 <= F2_MAX_DW && h->cells[c].ch <= F2_MAX_CH) maxch = std::max(maxch, (int)h->cells[c].ch);
            if (maxch > 6 && !getenv("YGZ_NO_TMA") && !getenv("YGZ_FAST_V1") && L.w >= F2_PITCH && L.h >= maxch + 1 &&
                make_tensor_map_u8_3d(&h->fastmaps.level[l], L.d_base, L.w, L.h, B, L.pitch, L.frame_stride, F2_PITCH, maxch + 1) == 0) {
                ok |= 1u << l;
                h->fastmaps.boxh[l] = (unsigned char)(maxch + 1);
            }
        }
        for (int c = 0; c < (int)h->cells.size(); c++) {
            const CellDesc& cd = h->cells[c];
            const bool v2 = ((ok >> cd.level) & 1u) && cd.cw > 6 && cd.ch > 6 && cd.cw - 6 <= F2_MAX_DW && cd.ch <= F2_MAX_CH;
            (v2 ? narrow : wide).push_back(c);
        }
        h->n_fast_narrow = (int)narrow.size();
        h->n_fast_wide = (int)wide.size();
        narrow.insert(narrow.end(), wide.begin(), wide.end());
        TRY_OR_FREE(cudaMalloc(&h->d_fast_ids, std::max<size_t>(narrow.size(), 1) * sizeof(int)));
        if (!narrow.empty())
            TRY_OR_FREE(cudaMemcpy(h->d_fast_ids, narrow.data(), narrow.size() * sizeof(int), cudaMemcpyHostToDevice));
        TRY_OR_FREE(cudaMalloc(&h->d_fast_queue, 64));
        TRY_OR_FREE(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
        TRY_OR_FREE(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        TRY_OR_FREE(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
        TRY_OR_FREE(cudaFuncSetAttribute(k_fast_cells_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_WARPS * F2_WARP_SMEM));
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        h->fast_v2_grid = sms * 4;                       // persistent: 4 resident CTAs per SM (registers and shared memory)
        if (const char* e = getenv("YGZ_FAST_V2_CTAS")) h->fast_v2_grid = sms * std::max(1, atoi(e));
    }
    {
        std::vector<float> fp(256 * 4);
        for (int t = 0; t < 256; t++)            // test t = 8 * byte + j  ->  slot j * 32 + byte
            for (int c = 0; c < 4; c++) fp[((t & 7) * 32 + (t >> 3)) * 4 + c] = (float)h_pattern[t * 4 + c];
        TRY_OR_FREE(cudaMalloc(&h->d_pattern, fp.size() * sizeof(float)));
        TRY_OR_FREE(cudaMemcpy(h->d_pattern, fp.data(), fp.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    TRY_OR_FREE(cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(fast_smem, 1024)));
    TRY_OR_FREE(cudaFuncSetAttribute(k_octree<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(oct_smem, 1024)));
    TRY_OR_FREE(cudaFuncSetAttribute(k_octree<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(oct_smem, 1024)));
    TRY_OR_FREE(cudaDeviceSynchronize());
#undef TRY_OR_FREE
    *out = h;
    return YGZ_OK;
}

int ygz_orb_destroy(ygz_orb* h) { return orb_free(h); }

int ygz_orb_tables(ygz_orb* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* fpl) {
    YGZ_ARG_CHECK(h);
    for (int i = 0; i < h->cfg.nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->inv_scale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
        if (fpl) fpl[i] = h->feat_per_level[i];
    }
    return YGZ_OK;
}

int ygz_orb_level_size(ygz_orb* h, int level, int* w, int* hgt) {
    YGZ_ARG_CHECK(h && level >= 0 && level < h->cfg.nlevels);
    if (w) *w = h->lv[level].w;
    if (hgt) *hgt = h->lv[level].h;
    return YGZ_OK;
}

int ygz_orb_max_keypoints(ygz_orb* h) { return h ? h->P.out_cap : YGZ_ERR_ARG; }
int ygz_orb_max_batch(ygz_orb* h) { return h ? h->cfg.max_batch : YGZ_ERR_ARG; }
void* ygz_orb_stream(ygz_orb* h) { return h ? (void*)h->stream : nullptr; }

int ygz_orb_upload(ygz_orb* h, const uint8_t* images, int n, size_t frame_stride, int row_stride) {
    YGZ_ARG_CHECK(h && images && n >= 1 && n <= h->cfg.max_batch && row_stride >= h->cfg.width);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    const LevelHost& L0 = h->lv[0];
    if (frame_stride == (size_t)row_stride * L0.h && row_stride == L0.pitch) {
        // contiguous frames and no row padding on the device (w % 16 == 0): one 1-D copy
        YGZ_CUDA_TRY(cudaMemcpyAsync(L0.d_base, images, L0.frame_stride * n, cudaMemcpyHostToDevice, h->stream));
    } else if (frame_stride == (size_t)row_stride * L0.h) {
        // Contiguous frames whose row length is not the device pitch (w % 16 != 0, e.g. 1242): a 2-D host-to-device copy of
        // n*h short rows runs at a fraction of the link rate (measured 10 GB/s on 1242-byte rows), so the block crosses PCIe as
        // ONE 1-D copy into a staging buffer and is re-pitched on the device.
        const size_t bytes = frame_stride * n;
        if (bytes > h->stage_bytes) {
            cudaFree(h->d_stage); h->d_stage = nullptr; h->stage_bytes = 0;
            YGZ_CUDA_TRY(cudaMalloc(&h->d_stage, frame_stride * h->cfg.max_batch));
            h->stage_bytes = frame_stride * h->cfg.max_batch;
        }
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_stage, images, bytes, cudaMemcpyHostToDevice, h->stream));
        YGZ_CUDA_TRY(cudaMemcpy2DAsync(L0.d_base, L0.pitch, h->d_stage, row_stride, L0.w, (size_t)L0.h * n,
                                       cudaMemcpyDeviceToDevice, h->stream));
    } else {
        for (int f = 0; f < n; f++)
            YGZ_CUDA_TRY(cudaMemcpy2DAsync(L0.d_base + f * L0.frame_stride, L0.pitch, images + f * frame_stride, row_stride,
                                           L0.w, L0.h, cudaMemcpyHostToDevice, h->stream));
    }
    h->ctr.bytes_in += (uint64_t)L0.w * L0.h * n;
    h->last_n = n;
    return YGZ_OK;
}

// kernels launch_detect issues: k_fast_cells_v2 for the narrow cells, k_fast_cells for the rest
static inline int fast_launches(const ygz_orb* h) { return (h->n_fast_narrow > 0) + (h->n_fast_wide > 0); }

static int launch_pyramid(ygz_orb* h, int n) {
    NvtxRange nvtx_range("ygz/ComputePyramid");
    for (int l = 1; l < h->cfg.nlevels; l++) {
        const LevelHost& S = h->lv[l - 1];
        const LevelHost& D = h->lv[l];
        if (D.py2_smem > 0) {
            dim3 grid(ceil_div(D.w, PY2_COLS), ceil_div(D.h, PY2_TILE_ROWS), n);
            k_pyr_resize_v2<<<grid, PY2_THREADS, D.py2_smem, h->stream>>>(S.d_base, S.pitch, S.frame_stride, S.w, S.h, D.d_base, D.pitch,
                                                                          D.frame_stride, D.w, D.h, D.d_tx, D.d_ty, D.py2_pitch);
            continue;
        }
        dim3 block(PYR_THREADS);
        dim3 grid(ceil_div(ceil_div(D.w, 4), PYR_THREADS), ceil_div(D.h, PYR_ROWS), n);
        k_pyr_resize<<<grid, block, 0, h->stream>>>(S.d_base, S.pitch, S.frame_stride, S.w, S.h, D.d_base, D.pitch, D.frame_stride,
                                                    D.w, D.h, D.d_tx, D.d_ty);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    return YGZ_OK;
}

static int launch_detect(ygz_orb* h, int n) {
    NvtxRange nvtx_range("ygz/FAST cells");
    // the few cells the TMA kernel does not take go first, on the auxiliary stream: one cell per warp is a latency-bound
    // launch (55 us for 1.5 % of the C1 cells when serialised) that hides completely beside the persistent kernel
    const bool fork = h->n_fast_wide > 0 && h->n_fast_narrow > 0;
    if (h->n_fast_wide > 0) {
        cudaStream_t s = h->stream;
        if (fork) {
            YGZ_CUDA_TRY(cudaEventRecord(h->ev_fork, h->stream));
            YGZ_CUDA_TRY(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
            s = h->aux_stream;
        }
        dim3 grid(ceil_div(h->n_fast_wide, FAST_WARPS), n);
        k_fast_cells<<<grid, FAST_THREADS, h->fast_smem, s>>>(h->P, h->d_cells, h->d_fast_ids + h->n_fast_narrow, h->n_fast_wide,
                                                              h->d_pool, h->d_cell_count, h->fast_warp_smem);
        if (fork) YGZ_CUDA_TRY(cudaEventRecord(h->ev_join, h->aux_stream));
    }
    if (h->n_fast_narrow > 0) {
        const int items = h->n_fast_narrow * n;
        const int grid = std::min(h->fast_v2_grid, ceil_div(items, F2_WARPS));
        // items are drawn from the queue (one atomic per cell); a static share can be dealt round-robin first
        static const int static_pct = getenv("YGZ_FAST_STATIC_PCT") ? atoi(getenv("YGZ_FAST_STATIC_PCT")) : 0;   // tuning knob; measured on B200 (C1, B=256, 4 CTAs/SM): 0 % 0.770 ms, 50 % 0.780, 75 % 0.818, 90 % 0.835
        const int n_static = std::min(items, (int)((long long)items * static_pct / 100) / (grid * F2_WARPS) * (grid * F2_WARPS));
        YGZ_CUDA_TRY(cudaMemsetAsync(h->d_fast_queue, 0, 4, h->stream));
        k_fast_cells_v2<<<grid, F2_THREADS, F2_WARPS * F2_WARP_SMEM, h->stream>>>(h->P, h->fastmaps, h->d_cells, h->d_fast_ids, h->n_fast_narrow, n,
                                                                                  n_static, h->d_fast_queue, h->d_pool, h->d_cell_count);
        if (fork) YGZ_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    return YGZ_OK;
}

static int launch_octree(ygz_orb* h, int n) {
    NvtxRange nvtx_range("ygz/DistributeOctTree");
    dim3 grid(n, h->cfg.nlevels);
    if (h->oct_unroll == 4)
        k_octree<4><<<grid, OCT_THREADS, h->oct_smem, h->stream>>>(h->P, h->d_cells, h->d_pool, h->d_cell_count, h->d_keys,
                                                                   h->d_keynode, h->d_nkeys, h->d_sel, h->d_nsel, h->oct_node_smem, h->oct_key_cache);
    else
        k_octree<8><<<grid, OCT_THREADS, h->oct_smem, h->stream>>>(h->P, h->d_cells, h->d_pool, h->d_cell_count, h->d_keys,
                                                                   h->d_keynode, h->d_nkeys, h->d_sel, h->d_nsel, h->oct_node_smem, h->oct_key_cache);
    YGZ_CUDA_TRY(cudaGetLastError());
    return YGZ_OK;
}

static int launch_blur(ygz_orb* h, int n) {
    NvtxRange nvtx_range("ygz/GaussianBlur");
    dim3 grid(h->n_blur_tiles, n);
    k_blur_levels<<<grid, BLUR_THREADS, 0, h->stream>>>(h->P, h->tma, h->d_blur_tiles);
    YGZ_CUDA_TRY(cudaGetLastError());
    return YGZ_OK;
}

static int launch_describe(ygz_orb* h, int n, int n_existing) {
    NvtxRange nvtx_range("ygz/IC_Angle+rBRIEF");
    dim3 grid(ceil_div(h->P.sel_total, DESC_WARPS), n);
    k_orient_describe<<<grid, DESC_THREADS, 0, h->stream>>>(h->P, h->patch, 0, n_existing, h->P.sel_total, h->d_pattern, h->d_sel, h->d_nsel, nullptr,
                                                           h->d_kps, h->d_desc, h->d_counts);
    if (n_existing > 0) {
        dim3 g2(ceil_div(n_existing, DESC_WARPS), n);
        k_orient_describe<<<g2, DESC_THREADS, 0, h->stream>>>(h->P, h->patch, 1, n_existing, n_existing, h->d_pattern, h->d_sel, h->d_nsel, h->d_existing,
                                                             h->d_kps, h->d_desc, h->d_counts);
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    return YGZ_OK;
}

int ygz_orb_run(ygz_orb* h, int n) {
    YGZ_ARG_CHECK(h && n >= 1 && n <= h->cfg.max_batch);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    int rc;
    if ((rc = launch_pyramid(h, n))) return rc;
    if ((rc = launch_detect(h, n))) return rc;
    // (measured, C2 B=256: GaussianBlur on the auxiliary stream beside DistributeOctTree changes nothing — 126.5 vs 126.4 k frames/s;
    //  k_octree's 4 CTAs/SM hold the whole register file, so the blur CTAs only ever fill its tail, which the second batch in flight does already)
    if ((rc = launch_octree(h, n))) return rc;
    if ((rc = launch_blur(h, n))) return rc;
    if ((rc = launch_describe(h, n, 0))) return rc;
    h->launches_per_run = (h->cfg.nlevels - 1) + 3 + fast_launches(h);
    h->ctr.kernel_launches += h->launches_per_run;
    h->ctr.frames += n;
    h->last_n = n;
    return YGZ_OK;
}

int ygz_orb_sync(ygz_orb* h) {
    YGZ_ARG_CHECK(h);
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return YGZ_OK;
}

int ygz_orb_download(ygz_orb* h, int n, ygz_keypoint* kps, uint8_t* desc, int* counts, int cap) {
    YGZ_ARG_CHECK(h && n >= 1 && n <= h->cfg.max_batch && counts && cap >= 0);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    const int oc = h->P.out_cap;
    YGZ_CUDA_TRY(cudaMemcpyAsync(counts, h->d_counts, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    const int m = std::min(cap, oc);
    if (kps && m > 0)
        YGZ_CUDA_TRY(cudaMemcpy2DAsync(kps, (size_t)cap * sizeof(ygz_keypoint), h->d_kps, (size_t)oc * sizeof(ygz_keypoint),
                                       (size_t)m * sizeof(ygz_keypoint), n, cudaMemcpyDeviceToHost, h->stream));
    if (desc && m > 0)
        YGZ_CUDA_TRY(cudaMemcpy2DAsync(desc, (size_t)cap * 32, h->d_desc, (size_t)oc * 32, (size_t)m * 32, n,
                                       cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    uint64_t tot = 0;
    int rc = YGZ_OK;
    for (int f = 0; f < n; f++) {
        tot += counts[f];
        if (counts[f] > cap) rc = YGZ_ERR_CAPACITY;
    }
    h->ctr.keypoints += tot;
    h->ctr.bytes_out += (uint64_t)n * 4 + (uint64_t)n * m * (sizeof(ygz_keypoint) + 32);
    if (rc) set_last_error("keypoint buffer too small (cap < count)");
    return rc;
}

int ygz_orb_extract(ygz_orb* h, const uint8_t* images, int n, size_t frame_stride, int row_stride, ygz_keypoint* kps,
                    uint8_t* desc, int* counts, int cap) {
    int rc;
    if (!images || n <= 0) {            // operator() returns silently on an empty image (src/ORBextractor.cc:972-973)
        if (counts) for (int i = 0; i < n; i++) counts[i] = 0;
        return YGZ_OK;
    }
    if ((rc = ygz_orb_upload(h, images, n, frame_stride, row_stride))) return rc;
    if ((rc = ygz_orb_run(h, n))) return rc;
    return ygz_orb_download(h, n, kps, desc, counts, cap);
}

// Asynchronous variant of ygz_orb_extract: enqueues H2D, all kernels and the D2H of the results on the handle's stream
// and returns; the (pinned) host buffers must stay valid until ygz_orb_wait.  Two handles used alternately overlap
// one batch's copies with the other's kernels.
int ygz_orb_extract_async(ygz_orb* h, const uint8_t* images, int n, size_t frame_stride, int row_stride, ygz_keypoint* kps,
                          uint8_t* desc, int* counts, int cap) {
    int rc;
    YGZ_ARG_CHECK(h && images && n >= 1 && kps && desc && counts);
    if ((rc = ygz_orb_upload(h, images, n, frame_stride, row_stride))) return rc;
    if ((rc = ygz_orb_run(h, n))) return rc;
    const int oc = h->P.out_cap, m = std::min(cap, oc);
    YGZ_CUDA_TRY(cudaMemcpyAsync(counts, h->d_counts, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    if (m > 0) {
        YGZ_CUDA_TRY(cudaMemcpy2DAsync(kps, (size_t)cap * sizeof(ygz_keypoint), h->d_kps, (size_t)oc * sizeof(ygz_keypoint),
                                       (size_t)m * sizeof(ygz_keypoint), n, cudaMemcpyDeviceToHost, h->stream));
        YGZ_CUDA_TRY(cudaMemcpy2DAsync(desc, (size_t)cap * 32, h->d_desc, (size_t)oc * 32, (size_t)m * 32, n, cudaMemcpyDeviceToHost, h->stream));
    }
    h->ctr.bytes_out += (uint64_t)n * 4 + (uint64_t)n * m * (sizeof(ygz_keypoint) + 32);
    return YGZ_OK;
}

int ygz_orb_wait(ygz_orb* h) {
    YGZ_ARG_CHECK(h);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return YGZ_OK;
}

int ygz_orb_compute_pyramid(ygz_orb* h, const uint8_t* images, int n, size_t frame_stride, int row_stride) {
    int rc;
    if ((rc = ygz_orb_upload(h, images, n, frame_stride, row_stride))) return rc;
    if ((rc = launch_pyramid(h, n))) return rc;
    h->ctr.kernel_launches += h->cfg.nlevels - 1;
    return ygz_orb_sync(h);
}

int ygz_orb_extract_frame(ygz_orb* h, const uint8_t* image, int row_stride, int n_existing, ygz_keypoint* kps, uint8_t* desc,
                          int* count, int cap) {
    YGZ_ARG_CHECK(h && image && kps && desc && count && n_existing >= 0 && cap >= n_existing);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    int rc;
    if ((rc = ygz_orb_upload(h, image, 1, (size_t)row_stride * h->cfg.height, row_stride))) return rc;
    if (n_existing > h->existing_cap) {
        cudaFree(h->d_existing);
        h->d_existing = nullptr;
        // outputs need room for the existing keys too
        YGZ_CUDA_TRY(cudaMalloc(&h->d_existing, (size_t)n_existing * sizeof(ExistingKey)));
        h->existing_cap = n_existing;
    }
    if (n_existing + h->P.sel_total > h->P.out_cap) {
        // grow the output arrays so existing + new keys fit
        const int B = h->cfg.max_batch, oc = n_existing + h->P.sel_total;
        cudaFree(h->d_kps); cudaFree(h->d_desc);
        h->d_kps = nullptr; h->d_desc = nullptr;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_kps, (size_t)oc * B * sizeof(ygz_keypoint)));
        YGZ_CUDA_TRY(cudaMalloc(&h->d_desc, (size_t)oc * B * 32));
        h->P.out_cap = oc;
    }
    std::vector<ExistingKey> ek(n_existing);
    for (int i = 0; i < n_existing; i++) {
        int oct = kps[i].octave;
        if (oct < 0 || oct >= h->cfg.nlevels) { set_last_error("existing key octave out of range"); return YGZ_ERR_ARG; }
        ek[i].x = kps[i].x * h->inv_scale[oct];        // tmp.pt *= mvInvScaleFactor[tmp.octave], :1095
        ek[i].y = kps[i].y * h->inv_scale[oct];
        ek[i].angle = kps[i].angle;
        ek[i].octave = oct;
    }
    if (n_existing)
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_existing, ek.data(), (size_t)n_existing * sizeof(ExistingKey), cudaMemcpyHostToDevice, h->stream));
    if ((rc = launch_pyramid(h, 1))) return rc;
    if ((rc = launch_detect(h, 1))) return rc;
    if ((rc = launch_octree(h, 1))) return rc;
    if ((rc = launch_blur(h, 1))) return rc;
    if ((rc = launch_describe(h, 1, n_existing))) return rc;
    h->ctr.kernel_launches += (h->cfg.nlevels - 1) + 3 + fast_launches(h) + (n_existing > 0);
    h->ctr.frames += 1;
    YGZ_CUDA_TRY(cudaMemcpyAsync(count, h->d_counts, 4, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (*count > cap) { set_last_error("keypoint buffer too small"); return YGZ_ERR_CAPACITY; }
    const int n_new = *count - n_existing;
    if (n_new > 0)
        YGZ_CUDA_TRY(cudaMemcpyAsync(kps + n_existing, h->d_kps + n_existing, (size_t)n_new * sizeof(ygz_keypoint),
                                     cudaMemcpyDeviceToHost, h->stream));
    if (*count > 0)
        YGZ_CUDA_TRY(cudaMemcpyAsync(desc, h->d_desc, (size_t)(*count) * 32, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->ctr.keypoints += *count;
    return YGZ_OK;
}

// scratch of the DSO selection for `frames` frames: occupancy maps, per-cell winners, compacted key lists, loop state
static int dso_ensure(ygz_orb* h, int frames) {
    if (frames <= h->dso_frames) return YGZ_OK;
    const int W = h->cfg.width, H = h->cfg.height;
    const size_t max_cells = (size_t)std::max((H / 7) * (W / 7), 1), F = (size_t)frames;
    cudaFree(h->d_occ); cudaFree(h->d_cell_keys); cudaFree(h->d_cell_cnt); cudaFree(h->d_dso_keys); cudaFree(h->d_dso_total); cudaFree(h->d_dso_state);
    h->d_occ = nullptr; h->d_cell_keys = nullptr; h->d_cell_cnt = nullptr; h->d_dso_keys = nullptr; h->d_dso_total = nullptr; h->d_dso_state = nullptr;
    h->dso_frames = 0;
    YGZ_CUDA_TRY(cudaMalloc(&h->d_occ, F * W * H));
    YGZ_CUDA_TRY(cudaMalloc(&h->d_cell_keys, F * max_cells * 3 * 4));
    YGZ_CUDA_TRY(cudaMalloc(&h->d_cell_cnt, F * max_cells));
    YGZ_CUDA_TRY(cudaMalloc(&h->d_dso_keys, F * max_cells * 3 * 4));
    YGZ_CUDA_TRY(cudaMalloc(&h->d_dso_total, 4));
    YGZ_CUDA_TRY(cudaMalloc(&h->d_dso_state, F * sizeof(DsoFrameState)));
    YGZ_CUDA_TRY(cudaFuncSetAttribute(k_dso_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, DSO_WARPS * dso_warp_smem(DSO_GMAX)));
    YGZ_CUDA_TRY(cudaFuncSetAttribute(k_dso_cells_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, DSO_WARPS * dso_warp_smem(DSO_GMAX)));
    h->dso_frames = frames;
    return YGZ_OK;
}

// operator()(Frame*, keypoints, descriptors, DSO_KEYPOINT), src/ORBextractor.cc:1031-1127: ComputeKeyPointsDSOSingleLevel
// (:1277-1381) on level 0 + descriptors of the pre-existing and the new keys.  One frame per call: the adaptive cell size
// mnGridSize is state carried from frame to frame (*grid_size in/out; NULL = the handle keeps it).
int ygz_orb_extract_frame_dso(ygz_orb* h, const uint8_t* image, int row_stride, int n_existing, ygz_keypoint* kps, uint8_t* desc,
                              int* count, int cap, int* grid_size) {
    YGZ_ARG_CHECK(h && image && kps && desc && count && n_existing >= 0 && cap >= n_existing);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    const int W = h->cfg.width, H = h->cfg.height, n = h->cfg.nfeatures;
    const LevelHost& L0 = h->lv[0];
    int rc;
    if ((rc = dso_ensure(h, 1))) return rc;
    if ((rc = ygz_orb_upload(h, image, 1, (size_t)row_stride * H, row_stride))) return rc;
    if (n_existing > h->existing_cap) {
        cudaFree(h->d_existing); h->d_existing = nullptr;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_existing, (size_t)n_existing * sizeof(ExistingKey)));
        h->existing_cap = n_existing;
    }
    if (n_existing > h->dso_pts_cap) {
        cudaFree(h->d_dso_pts); h->d_dso_pts = nullptr;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_dso_pts, (size_t)n_existing * sizeof(float2)));
        h->dso_pts_cap = n_existing;
    }
    std::vector<ExistingKey> ek(n_existing);
    std::vector<float2> pts(n_existing);
    for (int i = 0; i < n_existing; i++) {
        const int oct = kps[i].octave;
        if (oct < 0 || oct >= h->cfg.nlevels) { set_last_error("existing key octave out of range"); return YGZ_ERR_ARG; }
        ek[i].x = kps[i].x * h->inv_scale[oct];        // kp.pt * mvInvScaleFactor[kp.octave], :1095 and :1379
        ek[i].y = kps[i].y * h->inv_scale[oct];
        ek[i].angle = kps[i].angle;
        ek[i].octave = oct;
        pts[i] = make_float2(kps[i].x, kps[i].y);
    }
    YGZ_CUDA_TRY(cudaMemsetAsync(h->d_occ, 0, (size_t)W * H, h->stream));
    if (n_existing) {
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_existing, ek.data(), (size_t)n_existing * sizeof(ExistingKey), cudaMemcpyHostToDevice, h->stream));
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_dso_pts, pts.data(), (size_t)n_existing * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
        k_dso_mark_occupancy<<<ceil_div(n_existing, 256), 256, 0, h->stream>>>(h->d_dso_pts, n_existing, h->d_occ, W, H);
    }
    if ((rc = launch_pyramid(h, 1))) return rc;
    if ((rc = launch_blur(h, 1))) return rc;
    int launches = (h->cfg.nlevels - 1) + 1 + (n_existing > 0);
    // ---- the while (cnt < n) loop of :1300-1371, one device pass per cell size
    int g = grid_size ? *grid_size : h->dso_grid;
    if (g < 0) g = static_cast<int>(std::sqrt(1.0 * H * W / n));
    int cnt = 0;
    while (cnt < n) {
        if (cnt > 0) {
            g -= 5;
            if (g < 7) { g = 7; break; }
        }
        if (g < 7 || g > DSO_GMAX) { set_last_error("DSO grid size outside [7, 96]"); return YGZ_ERR_ARG; }
        DsoParams D;
        D.img = L0.d_base; D.occ = h->d_occ; D.w = W; D.h = H; D.pitch = L0.pitch;
        D.g = g; D.n_rows = H / g; D.n_cols = W / g;
        const int cells = D.n_rows * D.n_cols;
        cnt = 0;
        if (cells > 0) {
            const int ws = dso_warp_smem(g);
            k_dso_cells<<<ceil_div(cells, DSO_WARPS), DSO_WARPS * 32, (size_t)DSO_WARPS * ws, h->stream>>>(D, h->d_cell_keys, h->d_cell_cnt, ws);
            k_dso_compact<<<1, 1024, 0, h->stream>>>(h->d_cell_keys, h->d_cell_cnt, cells, h->d_dso_keys, h->d_dso_total);
            YGZ_CUDA_TRY(cudaGetLastError());
            YGZ_CUDA_TRY(cudaMemcpyAsync(&cnt, h->d_dso_total, 4, cudaMemcpyDeviceToHost, h->stream));
            YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
            launches += 2;
        }
        if (cnt == 0) break;       // the reference would spin forever on a frame without corners; the oracle stops here
    }
    if (cnt > n) g += 5;           // :1373-1375
    if (grid_size) *grid_size = g;
    h->dso_grid = g;
    // ---- descriptors: existing keys (angles refreshed, :1377-1380) then the new level-0 keys
    if (n_existing + cnt > h->P.out_cap) {
        const int B = h->cfg.max_batch, oc = n_existing + cnt;
        cudaFree(h->d_kps); cudaFree(h->d_desc);
        h->d_kps = nullptr; h->d_desc = nullptr;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_kps, (size_t)oc * B * sizeof(ygz_keypoint)));
        YGZ_CUDA_TRY(cudaMalloc(&h->d_desc, (size_t)oc * B * 32));
        h->P.out_cap = oc;
    }
    *count = n_existing + cnt;
    if (*count > cap) { set_last_error("keypoint buffer too small"); return YGZ_ERR_CAPACITY; }
    if (n_existing > 0) {
        k_orient_describe<<<dim3(ceil_div(n_existing, DESC_WARPS), 1), DESC_THREADS, 0, h->stream>>>(
            h->P, h->patch, 3, n_existing, n_existing, h->d_pattern, h->d_sel, h->d_nsel, h->d_existing, h->d_kps, h->d_desc, h->d_counts);
        launches++;
    }
    if (cnt > 0) {
        k_orient_describe<<<dim3(ceil_div(cnt, DESC_WARPS), 1), DESC_THREADS, 0, h->stream>>>(
            h->P, h->patch, 2, n_existing, cnt, h->d_pattern, h->d_dso_keys, h->d_nsel, nullptr, h->d_kps, h->d_desc, h->d_counts);
        launches++;
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    if (*count > 0) {
        std::vector<ygz_keypoint> tmp((size_t)*count);
        YGZ_CUDA_TRY(cudaMemcpyAsync(tmp.data(), h->d_kps, (size_t)(*count) * sizeof(ygz_keypoint), cudaMemcpyDeviceToHost, h->stream));
        YGZ_CUDA_TRY(cudaMemcpyAsync(desc, h->d_desc, (size_t)(*count) * 32, cudaMemcpyDeviceToHost, h->stream));
        YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int i = 0; i < n_existing; i++) kps[i].angle = tmp[i].angle;
        for (int i = n_existing; i < *count; i++) kps[i] = tmp[i];
    }
    h->ctr.kernel_launches += launches;
    h->ctr.frames += 1;
    h->ctr.keypoints += *count;
    return YGZ_OK;
}

// The same for n_frames frames of n_frames INDEPENDENT streams (one mnGridSize each: grid_sizes[f] in / out, -1 before a stream's first
// frame).  kps / desc are [n_frames][cap]; the first n_existing[f] key points of row f are that frame's direct-tracked keys.
// The adaptive loop runs on the device (k_dso_cells_batch / k_dso_compact_batch, four passes per host synchronisation).
int ygz_orb_extract_frame_dso_batch(ygz_orb* h, const uint8_t* images, int n_frames, size_t frame_stride, int row_stride, const int* n_existing,
                                    ygz_keypoint* kps, uint8_t* desc, int* counts, int cap, int* grid_sizes) {
    YGZ_ARG_CHECK(h && images && n_frames >= 1 && n_frames <= h->cfg.max_batch && n_existing && kps && desc && counts && grid_sizes && cap >= 0);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    const int W = h->cfg.width, H = h->cfg.height, n = h->cfg.nfeatures, B = n_frames;
    const LevelHost& L0 = h->lv[0];
    const int max_cells = std::max((H / 7) * (W / 7), 1);
    int rc;
    if ((rc = dso_ensure(h, B))) return rc;
    int ex_stride = 1, ex_total = 0;
    for (int f = 0; f < B; f++) {
        YGZ_ARG_CHECK(n_existing[f] >= 0 && n_existing[f] <= cap);
        ex_stride = std::max(ex_stride, n_existing[f]);
        ex_total += n_existing[f];
    }
    if ((rc = ygz_orb_upload(h, images, B, frame_stride, row_stride))) return rc;
    if (B * ex_stride > h->existing_cap) {
        cudaFree(h->d_existing); h->d_existing = nullptr; h->existing_cap = 0;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_existing, (size_t)B * ex_stride * sizeof(ExistingKey)));
        h->existing_cap = B * ex_stride;
    }
    if (ex_total > h->dso_pts_cap) {
        cudaFree(h->d_dso_pts); h->d_dso_pts = nullptr; h->dso_pts_cap = 0;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_dso_pts, (size_t)ex_total * sizeof(float2)));
        h->dso_pts_cap = ex_total;
    }
    if (ex_total > h->dso_fid_cap) {
        cudaFree(h->d_dso_fid); h->d_dso_fid = nullptr; h->dso_fid_cap = 0;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_dso_fid, (size_t)ex_total * sizeof(int)));
        h->dso_fid_cap = ex_total;
    }
    std::vector<ExistingKey> ek((size_t)B * ex_stride);
    std::vector<float2> pts((size_t)ex_total);
    std::vector<int> fid((size_t)ex_total);
    std::vector<DsoFrameState> st(B);
    int gmax = 0, gmin = 1 << 30;
    for (int f = 0, o = 0; f < B; f++) {
        const ygz_keypoint* K = kps + (size_t)f * cap;
        for (int i = 0; i < n_existing[f]; i++, o++) {
            const int oct = K[i].octave;
            if (oct < 0 || oct >= h->cfg.nlevels) { set_last_error("existing key octave out of range"); return YGZ_ERR_ARG; }
            ExistingKey& e = ek[(size_t)f * ex_stride + i];
            e.x = K[i].x * h->inv_scale[oct];          // kp.pt * mvInvScaleFactor[kp.octave], :1095 and :1379
            e.y = K[i].y * h->inv_scale[oct];
            e.angle = K[i].angle;
            e.octave = oct;
            pts[o] = make_float2(K[i].x, K[i].y);
            fid[o] = f;
        }
        int g = grid_sizes[f];
        if (g < 0) g = static_cast<int>(std::sqrt(1.0 * H * W / n));
        if (g < 7 || g > DSO_GMAX) { set_last_error("DSO grid size outside [7, 96]"); return YGZ_ERR_ARG; }
        st[f] = DsoFrameState{g, 0, 0, n_existing[f]};
        gmax = std::max(gmax, g); gmin = std::min(gmin, g);
    }
    YGZ_CUDA_TRY(cudaMemsetAsync(h->d_occ, 0, (size_t)B * W * H, h->stream));
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_dso_state, st.data(), (size_t)B * sizeof(DsoFrameState), cudaMemcpyHostToDevice, h->stream));
    if (ex_total) {
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_existing, ek.data(), ek.size() * sizeof(ExistingKey), cudaMemcpyHostToDevice, h->stream));
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_dso_pts, pts.data(), (size_t)ex_total * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
        YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_dso_fid, fid.data(), (size_t)ex_total * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        k_dso_mark_occupancy_batch<<<ceil_div(ex_total, 256), 256, 0, h->stream>>>(h->d_dso_pts, h->d_dso_fid, ex_total, h->d_occ, W, H);
    }
    if ((rc = launch_pyramid(h, B))) return rc;
    if ((rc = launch_blur(h, B))) return rc;
    int launches = (h->cfg.nlevels - 1) + 1 + (ex_total > 0);
    DsoParams D;
    D.img = L0.d_base; D.occ = h->d_occ; D.w = W; D.h = H; D.pitch = L0.pitch; D.g = 0; D.n_rows = 0; D.n_cols = 0;
    const size_t smem = (size_t)DSO_WARPS * dso_warp_smem(gmax);          // cell sizes only shrink inside the loop
    int passes = 0;
    for (bool all_done = false; !all_done;) {
        for (int k = 0; k < 4; k++, passes++) {
            const int g_lo = std::max(7, gmin - 5 * passes);               // smallest cell size any frame can have in this pass
            const int cells = (H / g_lo) * (W / g_lo);
            if (cells > 0)
                k_dso_cells_batch<<<dim3(ceil_div(cells, DSO_WARPS), B), DSO_WARPS * 32, smem, h->stream>>>(D, L0.frame_stride, max_cells, h->d_dso_state,
                                                                                                           h->d_cell_keys, h->d_cell_cnt);
            k_dso_compact_batch<<<B, 1024, 0, h->stream>>>(h->d_cell_keys, h->d_cell_cnt, W, H, max_cells, n, h->d_dso_keys, h->d_dso_state);
            launches += 2;
        }
        YGZ_CUDA_TRY(cudaGetLastError());
        YGZ_CUDA_TRY(cudaMemcpyAsync(st.data(), h->d_dso_state, (size_t)B * sizeof(DsoFrameState), cudaMemcpyDeviceToHost, h->stream));
        YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
        all_done = true;
        for (int f = 0; f < B; f++) all_done = all_done && st[f].done;
        if (passes > 40) { set_last_error("DSO selection did not terminate"); return YGZ_ERR_CUDA; }
    }
    int max_cnt = 0, max_count = 0;
    for (int f = 0; f < B; f++) {
        counts[f] = n_existing[f] + st[f].cnt;                  // also on YGZ_ERR_CAPACITY: tells the caller what to allocate
        max_cnt = std::max(max_cnt, st[f].cnt);
        max_count = std::max(max_count, counts[f]);
    }
    if (max_count > cap) { set_last_error("keypoint buffer too small"); return YGZ_ERR_CAPACITY; }     // grid_sizes untouched: the call can be repeated
    for (int f = 0; f < B; f++) grid_sizes[f] = st[f].g;
    if (max_count > h->P.out_cap) {
        const int MB = h->cfg.max_batch;
        cudaFree(h->d_kps); cudaFree(h->d_desc);
        h->d_kps = nullptr; h->d_desc = nullptr;
        YGZ_CUDA_TRY(cudaMalloc(&h->d_kps, (size_t)max_count * MB * sizeof(ygz_keypoint)));
        YGZ_CUDA_TRY(cudaMalloc(&h->d_desc, (size_t)max_count * MB * 32));
        h->P.out_cap = max_count;
    }
    const int4* per_frame = reinterpret_cast<const int4*>(h->d_dso_state);
    if (ex_total > 0) {
        k_orient_describe<<<dim3(ceil_div(ex_stride, DESC_WARPS), B), DESC_THREADS, 0, h->stream>>>(
            h->P, h->patch, 3, ex_stride, ex_stride, h->d_pattern, h->d_sel, h->d_nsel, h->d_existing, h->d_kps, h->d_desc, h->d_counts, per_frame);
        launches++;
    }
    if (max_cnt > 0) {
        k_orient_describe<<<dim3(ceil_div(max_cnt, DESC_WARPS), B), DESC_THREADS, 0, h->stream>>>(
            h->P, h->patch, 2, ex_stride, max_cells * 3, h->d_pattern, h->d_dso_keys, h->d_nsel, nullptr, h->d_kps, h->d_desc, h->d_counts, per_frame);
        launches++;
    }
    YGZ_CUDA_TRY(cudaGetLastError());
    YGZ_CUDA_TRY(cudaMemcpyAsync(h->d_counts, counts, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, h->stream));     // zero-copy chaining sees the counts
    uint64_t total = 0;
    if (max_count > 0) {
        const int oc = h->P.out_cap;
        std::vector<ygz_keypoint> tmp((size_t)B * max_count);
        YGZ_CUDA_TRY(cudaMemcpy2DAsync(tmp.data(), (size_t)max_count * sizeof(ygz_keypoint), h->d_kps, (size_t)oc * sizeof(ygz_keypoint),
                                       (size_t)max_count * sizeof(ygz_keypoint), B, cudaMemcpyDeviceToHost, h->stream));
        YGZ_CUDA_TRY(cudaMemcpy2DAsync(desc, (size_t)cap * 32, h->d_desc, (size_t)oc * 32, (size_t)max_count * 32, B, cudaMemcpyDeviceToHost, h->stream));
        YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int f = 0; f < B; f++) {
            ygz_keypoint* K = kps + (size_t)f * cap;
            const ygz_keypoint* T = tmp.data() + (size_t)f * max_count;
            for (int i = 0; i < n_existing[f]; i++) K[i].angle = T[i].angle;
            for (int i = n_existing[f]; i < counts[f]; i++) K[i] = T[i];
            total += counts[f];
        }
    } else {
        YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    h->ctr.kernel_launches += launches;
    h->ctr.frames += B;
    h->ctr.keypoints += total;
    h->last_n = B;
    return YGZ_OK;
}

int ygz_orb_get_level(ygz_orb* h, int frame, int level, int border, uint8_t* out) {
    YGZ_ARG_CHECK(h && out && frame >= 0 && frame < h->cfg.max_batch && level >= 0 && level < h->cfg.nlevels && border >= 0);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    const LevelHost& L = h->lv[level];
    YGZ_ARG_CHECK(border < L.w && border < L.h);
    const int ow = L.w + 2 * border;
    YGZ_CUDA_TRY(cudaMemcpy2DAsync(out + (size_t)border * ow + border, ow, L.d_base + frame * L.frame_stride, L.pitch, L.w, L.h,
                                   cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (border > 0) {      // copyMakeBorder(..., BORDER_REFLECT_101), src/ORBextractor.cc:1141-1146 (host index arithmetic)
        auto refl = [](int p, int len) { if (p < 0) p = -p; if (p >= len) p = 2 * len - 2 - p; return p; };
        for (int y = -border; y < L.h + border; y++) {
            uint8_t* D = out + (size_t)(y + border) * ow;
            const uint8_t* S = out + (size_t)(refl(y, L.h) + border) * ow + border;
            if (y < 0 || y >= L.h) for (int x = 0; x < L.w; x++) D[x + border] = S[x];
        }
        for (int y = 0; y < L.h + 2 * border; y++) {
            uint8_t* D = out + (size_t)y * ow;
            for (int x = -border; x < 0; x++) D[x + border] = D[refl(x, L.w) + border];
            for (int x = L.w; x < L.w + border; x++) D[x + border] = D[refl(x, L.w) + border];
        }
    }
    return YGZ_OK;
}

/* GaussianBlur(mvImagePyramid[level], 7x7, sigma 2, BORDER_REFLECT_101) of frame `frame` as the last extraction left it on the
 * device (src/ORBextractor.cc:1009-1010 / :1079-1084): continuous w*h bytes.  Parity instrumentation for the stage tests. */
int ygz_orb_get_blurred_level(ygz_orb* h, int frame, int level, uint8_t* out) {
    YGZ_ARG_CHECK(h && out && frame >= 0 && frame < h->cfg.max_batch && level >= 0 && level < h->cfg.nlevels);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    const LevelHost& L = h->lv[level];
    YGZ_CUDA_TRY(cudaMemcpy2DAsync(out, L.w, L.d_blur + frame * L.frame_stride, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, h->stream));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return YGZ_OK;
}

int ygz_orb_get_candidates(ygz_orb* h, int frame, int level, ygz_keypoint* out, int cap) {
    YGZ_ARG_CHECK(h && frame >= 0 && frame < h->cfg.max_batch && level >= 0 && level < h->cfg.nlevels);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    YGZ_CUDA_TRY(cudaStreamSynchronize(h->stream));
    int n = 0;
    YGZ_CUDA_TRY(cudaMemcpy(&n, h->d_nkeys + (size_t)frame * h->cfg.nlevels + level, 4, cudaMemcpyDeviceToHost));
    if (n > cap) return -n - 100;
    std::vector<uint32_t> k(n);
    if (n) YGZ_CUDA_TRY(cudaMemcpy(k.data(), h->d_keys + (size_t)frame * h->P.key_total + h->P.lv[level].key_off, (size_t)n * 4,
                                   cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
        out[i].x = (float)(k[i] & 0xFFF);
        out[i].y = (float)((k[i] >> 12) & 0xFFF);
        out[i].size = 7.f; out[i].angle = -1.f;
        out[i].response = (float)(k[i] >> 24);
        out[i].octave = 0; out[i].class_id = -1;
    }
    return n;
}

int ygz_orb_time(ygz_orb* h, int n, int iters, float* stage_ms, int* launches) {
    YGZ_ARG_CHECK(h && n >= 1 && n <= h->cfg.max_batch && iters >= 1 && stage_ms);
    YGZ_CUDA_TRY(cudaSetDevice(h->device));
    float acc[6] = {0, 0, 0, 0, 0, 0};
    int rc;
    for (int it = 0; it < iters; it++) {
        YGZ_CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
        if ((rc = launch_pyramid(h, n))) return rc;
        YGZ_CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
        if ((rc = launch_detect(h, n))) return rc;
        YGZ_CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
        if ((rc = launch_octree(h, n))) return rc;
        YGZ_CUDA_TRY(cudaEventRecord(h->ev[3], h->stream));
        if ((rc = launch_blur(h, n))) return rc;
        YGZ_CUDA_TRY(cudaEventRecord(h->ev[4], h->stream));
        if ((rc = launch_describe(h, n, 0))) return rc;
        YGZ_CUDA_TRY(cudaEventRecord(h->ev[5], h->stream));
        YGZ_CUDA_TRY(cudaEventSynchronize(h->ev[5]));
        for (int s = 0; s < 5; s++) {
            float ms = 0;
            YGZ_CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[s], h->ev[s + 1]));
            acc[s] += ms;
        }
        float ms = 0;
        YGZ_CUDA_TRY(cudaEventElapsedTime(&ms, h->ev[0], h->ev[5]));
        acc[5] += ms;
    }
    for (int s = 0; s < 6; s++) stage_ms[s] = acc[s];
    if (launches) *launches = ((h->cfg.nlevels - 1) + 3 + fast_launches(h)) * iters;
    h->ctr.kernel_launches += (uint64_t)((h->cfg.nlevels - 1) + 3 + fast_launches(h)) * iters;
    h->ctr.frames += (uint64_t)n * iters;
    return YGZ_OK;
}

int ygz_orb_device_outputs(ygz_orb* h, void** d_kps, void** d_desc, void** d_counts, int* cap) {
    YGZ_ARG_CHECK(h);
    if (d_kps) *d_kps = h->d_kps;
    if (d_desc) *d_desc = h->d_desc;
    if (d_counts) *d_counts = h->d_counts;
    if (cap) *cap = h->P.out_cap;
    return YGZ_OK;
}

int ygz_orb_device_level(ygz_orb* h, int level, void** d_ptr, int* pitch, size_t* frame_stride) {
    YGZ_ARG_CHECK(h && level >= 0 && level < h->cfg.nlevels);
    if (d_ptr) *d_ptr = h->lv[level].d_base;
    if (pitch) *pitch = h->lv[level].pitch;
    if (frame_stride) *frame_stride = h->lv[level].frame_stride;
    return YGZ_OK;
}

namespace ygz_b200 {
__global__ void k_debug_cos_sin(unsigned first_bits, unsigned n, float* __restrict__ c, float* __restrict__ s) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float deg = __uint_as_float(first_bits + i);
    const float factorPI = (float)(3.141592653589793238462643383279502884197169399375 / 180.0);
    ref_cos_sin(__fmul_rn(deg, factorPI), &c[i], &s[i]);
}
}  // namespace ygz_b200

// (a, b) = (cos, sin) of computeOrbDescriptor (src/ORBextractor.cc:108-109) for the n consecutive float bit patterns
// starting at first_bits, interpreted as kpt.angle in degrees.  Parity instrumentation for tests/test_trig_gpu.py.
int ygz_debug_cos_sin(int device, unsigned first_bits, unsigned n, float* cos_out, float* sin_out) {
    YGZ_ARG_CHECK(cos_out && sin_out && n >= 1 && n <= (1u << 26));
    YGZ_CUDA_TRY(cudaSetDevice(device));
    float *dc = nullptr, *ds = nullptr;
    YGZ_CUDA_TRY(cudaMalloc(&dc, (size_t)n * 4));
    if (cudaMalloc(&ds, (size_t)n * 4) != cudaSuccess) { cudaFree(dc); set_last_error("cudaMalloc"); return YGZ_ERR_CUDA; }
    k_debug_cos_sin<<<(n + 255) / 256, 256>>>(first_bits, n, dc, ds);
    cudaError_t e = cudaMemcpy(cos_out, dc, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(sin_out, ds, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(dc); cudaFree(ds);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return YGZ_ERR_CUDA; }
    return YGZ_OK;
}

int ygz_orb_counters(ygz_orb* h, ygz_counters* out, int reset) {
    YGZ_ARG_CHECK(h && out);
    *out = h->ctr;
    if (reset) h->ctr = ygz_counters{};
    return YGZ_OK;
}

}  // extern "C"
