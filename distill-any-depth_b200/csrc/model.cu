// Host orchestration of the DepthAnythingV2 / DepthAnything forward (reference
// distillanydepth/depth_anything_v2/dpt.py:150-225, dinov2.py:212-321, util/blocks.py:29-148;
// teacher class modeling/archs/dam/dam.py:396-419 computes the same function).
//
// The model object owns an fp32 master copy of every parameter (student key layout) and, per
// precision mode, the packed operands the kernels read:
//   mode 0 (bf16)  bf16 K-major weight matrices for the tcgen05 GEMM / implicit-GEMM conv engine
//   mode 1 (fp32)  the same matrices in fp32 for the FFMA verification engine
// Activations live in a caller-provided workspace (bump-allocated, no cudaMalloc / sync in forward).
#include <cstdlib>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "backward.h"
#include "elementwise.h"
#include "gemm.h"

namespace dad {

struct ModelDesc {
    int embed_dim, depth, num_heads;
    int taps[4];
    int features;
    int out_channels[4];
};

namespace {

constexpr int PATCH_K = 588, PATCH_KP = 640, POS_GRID = 37;
constexpr float LN_EPS = 1e-6f;

__global__ void to_f32_kernel(const bf16* in, float* out, long long n) {
    const long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) out[i] = __bfloat162float(in[i]);
}

struct Mat {      // packed weight matrix, per mode
    void* w[2] = {nullptr, nullptr};
    int N = 0, Kp = 0;
};

struct Arena {    // bump allocator over the caller's workspace (dry run: only sizes)
    uint8_t* base;
    size_t cap, used = 0;
    bool dry;
    Arena(void* b, size_t c, bool d) : base(reinterpret_cast<uint8_t*>(b)), cap(c), dry(d) {}
    bool overflow = false;
    void* take(size_t bytes) {
        used = (used + 1023) & ~size_t(1023);
        void* p = dry ? nullptr : base + used;
        used += bytes;
        if (!dry && used > cap) { overflow = true; p = nullptr; }
        return p;
    }
};

}  // namespace

class Model {
public:
    explicit Model(const ModelDesc& d) : desc(d) {}
    ~Model() {
        for (auto& kv : master) cudaFree(kv.second.first);
        for (void* p : owned) cudaFree(p);
    }

    ModelDesc desc;
    std::unordered_map<std::string, std::pair<float*, long long>> master;  // name -> (fp32 device copy, numel)
    std::vector<void*> owned;
    bool packed[2] = {false, false};
    float* bqkv_scaled = nullptr;                       // [depth][3D] qkv bias with q part * 64^-0.5
    std::map<std::pair<int, int>, float*> pos_tables;   // (H, W) -> [1 + ph*pw, D]; nullptr = stale (weights changed)
    std::map<std::pair<int, int>, float*> pos_storage;  // the buffers behind pos_tables, reused across weight updates
    std::unordered_map<std::string, std::pair<float*, long long>> captures;
    std::unordered_map<std::string, std::pair<float*, long long>> grads;   // caller-owned fp32 gradient accumulators
    bool attn_bwd_fp32 = getenv("DAD_ATTN_BWD_FP32") != nullptr;   // A/B switch: bf16 training with the fp32 attention backward

    // packed matrices
    Mat patch;
    std::vector<Mat> qkv, proj, fc1, fc2;
    Mat projects[4], resize0, resize1, resize3, layer_rn[4];
    // optional parts, switched on by the PRESENCE of their parameters (no ABI field): the use_clstoken readout
    // (depth_head.readout_projects.{j}.0.*, dpt.py:116-122) and the ViT-g SwiGLU FFN (blocks.N.mlp.w12 / w3, swiglu_ffn.py)
    Mat readout_proj[4];
    std::vector<Mat> w12, w3;
    Mat rcu[4][2][2], out_conv[4], output_conv1, output_conv2_0;
    int CoP0 = 0, CoP1 = 0;

    int D() const { return desc.embed_dim; }
    bool has_readout() const { return master.count("depth_head.readout_projects.0.0.weight") != 0; }
    int swiglu_hidden() const {   // 0 = plain Mlp
        auto it = master.find("pretrained.blocks.0.mlp.w12.weight");
        return it == master.end() ? 0 : static_cast<int>(it->second.second / (2LL * desc.embed_dim));
    }

    const float* P(const std::string& name) const {
        auto it = master.find(name);
        return it == master.end() ? nullptr : it->second.first;
    }

    int set_weight(const char* name, const float* src, long long numel, cudaStream_t st) {
        DAD_REQUIRE(name && src && numel > 0, "set_weight: bad arguments");
        auto it = master.find(name);
        float* dst = nullptr;
        if (it != master.end()) {
            DAD_REQUIRE(it->second.second == numel, "set_weight: %s changed size (%lld -> %lld)", name,
                        it->second.second, numel);
            dst = it->second.first;
        } else {
            DAD_CHECK_CUDA(cudaMalloc(&dst, static_cast<size_t>(numel) * 4));
            master[name] = {dst, numel};
        }
        DAD_CHECK_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(numel) * 4, cudaMemcpyDeviceToDevice, st));
        packed[0] = packed[1] = false;  // repack lazily
        for (auto& kv : pos_tables) kv.second = nullptr;
        return DAD_OK;
    }

    int need(const std::string& name, long long numel) const {
        auto it = master.find(name);
        if (it == master.end()) return set_error(DAD_ERR_INVALID, "missing parameter %s", name.c_str());
        if (it->second.second != numel)
            return set_error(DAD_ERR_INVALID, "parameter %s has %lld elements, expected %lld", name.c_str(),
                             it->second.second, numel);
        return DAD_OK;
    }

    int alloc(void** p, size_t bytes) {
        DAD_CHECK_CUDA(cudaMalloc(p, bytes));
        owned.push_back(*p);
        return DAD_OK;
    }

    int pack_lin(Mat& m, int mode, const std::string& name, int N, int K, int Kp, int scale_rows, cudaStream_t st) {
        DAD_TRY(need(name, static_cast<long long>(N) * K));
        m.N = N; m.Kp = Kp;
        if (!m.w[mode]) DAD_TRY(alloc(&m.w[mode], static_cast<size_t>(N) * Kp * (mode == 0 ? 2 : 4)));
        return pack_linear(P(name), m.w[mode], mode == 0, N, K, Kp, scale_rows, 0.125f, st);
    }
    int pack_cv(Mat& m, int mode, const std::string& name, int Co, int Ci, int taps, cudaStream_t st) {
        DAD_TRY(need(name, static_cast<long long>(Co) * Ci * taps));
        const int Cp = cdiv(Ci, 64) * 64;
        m.N = Co; m.Kp = taps * Cp;
        if (!m.w[mode]) DAD_TRY(alloc(&m.w[mode], static_cast<size_t>(Co) * m.Kp * (mode == 0 ? 2 : 4)));
        return pack_conv(P(name), m.w[mode], mode == 0, Co, Ci, taps, Cp, st);
    }
    int pack_ct(Mat& m, int mode, const std::string& name, int C, int k, int& CoP, cudaStream_t st) {
        DAD_TRY(need(name, static_cast<long long>(C) * C * k * k));
        CoP = cdiv(C, 32) * 32;
        m.N = k * k * CoP; m.Kp = C;
        if (!m.w[mode]) DAD_TRY(alloc(&m.w[mode], static_cast<size_t>(m.N) * m.Kp * (mode == 0 ? 2 : 4)));
        return pack_convT(P(name), m.w[mode], mode == 0, C, C, k, CoP, m.Kp, st);
    }

    int pack(int mode, cudaStream_t st) {
        if (packed[mode]) return DAD_OK;
        const int Dm = D(), L = desc.depth, F = desc.features;
        const int* oc = desc.out_channels;
        DAD_REQUIRE(Dm % 64 == 0 && Dm == desc.num_heads * 64, "embed_dim must be heads*64");
        DAD_REQUIRE(F % 16 == 0, "features must be a multiple of 16");
        for (int j = 0; j < 4; ++j) DAD_REQUIRE(oc[j] % 8 == 0, "out_channels must be multiples of 8");
        const std::string p = "pretrained.";
        DAD_TRY(need(p + "cls_token", Dm));
        DAD_TRY(need(p + "pos_embed", static_cast<long long>(1 + POS_GRID * POS_GRID) * Dm));
        DAD_TRY(need(p + "patch_embed.proj.bias", Dm));
        DAD_TRY(need(p + "norm.weight", Dm));
        DAD_TRY(need(p + "norm.bias", Dm));
        DAD_TRY(pack_lin(patch, mode, p + "patch_embed.proj.weight", Dm, PATCH_K, PATCH_KP, 0, st));
        qkv.resize(L); proj.resize(L); fc1.resize(L); fc2.resize(L);
        // FFN flavour from the parameters present: Mlp (fc1 / fc2) or SwiGLUFFNFused (w12 / w3)
        const int ffn_hidden = swiglu_hidden();
        const bool swiglu_ffn = ffn_hidden > 0;
        if (swiglu_ffn) {
            DAD_REQUIRE(ffn_hidden % 8 == 0, "SwiGLU hidden width %d must be a multiple of 8", ffn_hidden);
            w12.resize(L); w3.resize(L);
        }
        if (!bqkv_scaled) DAD_TRY(alloc(reinterpret_cast<void**>(&bqkv_scaled), static_cast<size_t>(L) * 3 * Dm * 4));
        for (int i = 0; i < L; ++i) {
            const std::string b = p + "blocks." + std::to_string(i) + ".";
            for (const char* v : {"norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias", "ls1.gamma", "ls2.gamma",
                                  "attn.proj.bias"})
                DAD_TRY(need(b + v, Dm));
            DAD_TRY(need(b + "attn.qkv.bias", 3LL * Dm));
            DAD_TRY(pack_lin(qkv[i], mode, b + "attn.qkv.weight", 3 * Dm, Dm, Dm, Dm, st));  // q rows * 1/8
            DAD_TRY(copy_scale(P(b + "attn.qkv.bias"), bqkv_scaled + static_cast<long long>(i) * 3 * Dm, 3LL * Dm, Dm,
                               0.125f, st));
            DAD_TRY(pack_lin(proj[i], mode, b + "attn.proj.weight", Dm, Dm, Dm, 0, st));
            if (swiglu_ffn) {   // x12 = w12(x); w3(silu(x1) * x2)
                DAD_TRY(need(b + "mlp.w12.bias", 2LL * ffn_hidden));
                DAD_TRY(need(b + "mlp.w3.bias", Dm));
                DAD_TRY(pack_lin(w12[i], mode, b + "mlp.w12.weight", 2 * ffn_hidden, Dm, Dm, 0, st));
                DAD_TRY(pack_lin(w3[i], mode, b + "mlp.w3.weight", Dm, ffn_hidden, ffn_hidden, 0, st));
            } else {
                DAD_TRY(need(b + "mlp.fc1.bias", 4LL * Dm));
                DAD_TRY(need(b + "mlp.fc2.bias", Dm));
                DAD_TRY(pack_lin(fc1[i], mode, b + "mlp.fc1.weight", 4 * Dm, Dm, Dm, 0, st));
                DAD_TRY(pack_lin(fc2[i], mode, b + "mlp.fc2.weight", Dm, 4 * Dm, 4 * Dm, 0, st));
            }
        }
        const std::string h = "depth_head.";
        for (int j = 0; j < 4; ++j) {
            DAD_TRY(pack_lin(projects[j], mode, h + "projects." + std::to_string(j) + ".weight", oc[j], Dm, Dm, 0, st));
            DAD_TRY(need(h + "projects." + std::to_string(j) + ".bias", oc[j]));
        }
        for (int j = 0; has_readout() && j < 4; ++j) {   // Linear(2D -> D) + GELU on cat(patch tokens, cls)
            const std::string r = h + "readout_projects." + std::to_string(j) + ".0.";
            DAD_TRY(pack_lin(readout_proj[j], mode, r + "weight", Dm, 2 * Dm, 2 * Dm, 0, st));
            DAD_TRY(need(r + "bias", Dm));
        }
        DAD_TRY(pack_ct(resize0, mode, h + "resize_layers.0.weight", oc[0], 4, CoP0, st));
        DAD_TRY(pack_ct(resize1, mode, h + "resize_layers.1.weight", oc[1], 2, CoP1, st));
        DAD_TRY(pack_cv(resize3, mode, h + "resize_layers.3.weight", oc[3], oc[3], 9, st));
        DAD_TRY(need(h + "resize_layers.0.bias", oc[0]));
        DAD_TRY(need(h + "resize_layers.1.bias", oc[1]));
        DAD_TRY(need(h + "resize_layers.3.bias", oc[3]));
        const std::string s = h + "scratch.";
        for (int j = 0; j < 4; ++j)
            DAD_TRY(pack_cv(layer_rn[j], mode, s + "layer" + std::to_string(j + 1) + "_rn.weight", F, oc[j], 9, st));
        for (int r = 0; r < 4; ++r) {
            const std::string q = s + "refinenet" + std::to_string(r + 1) + ".";
            DAD_TRY(pack_cv(out_conv[r], mode, q + "out_conv.weight", F, F, 1, st));
            DAD_TRY(need(q + "out_conv.bias", F));
            for (int u = 0; u < 2; ++u)
                for (int c = 0; c < 2; ++c) {
                    const std::string n = q + "resConfUnit" + std::to_string(u + 1) + ".conv" + std::to_string(c + 1);
                    DAD_TRY(pack_cv(rcu[r][u][c], mode, n + ".weight", F, F, 9, st));
                    DAD_TRY(need(n + ".bias", F));
                }
        }
        DAD_TRY(pack_cv(output_conv1, mode, s + "output_conv1.weight", F / 2, F, 9, st));
        DAD_TRY(need(s + "output_conv1.bias", F / 2));
        DAD_TRY(pack_cv(output_conv2_0, mode, s + "output_conv2.0.weight", 32, F / 2, 9, st));
        DAD_TRY(need(s + "output_conv2.0.bias", 32));
        DAD_TRY(need(s + "output_conv2.2.weight", 32));
        DAD_TRY(need(s + "output_conv2.2.bias", 1));
        packed[mode] = true;
        return DAD_OK;
    }

    int prepare(int mode, int H, int W, cudaStream_t st) {
        DAD_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (bf16) or 1 (fp32)");
        DAD_REQUIRE(H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0, "H=%d, W=%d must be positive multiples of 14", H, W);
        DAD_TRY(pack(mode, st));
        float*& tab = pos_tables[std::make_pair(H, W)];
        const size_t rows = 1 + static_cast<size_t>(H / 14) * (W / 14);
        if (!tab) {
            float*& buf = pos_storage[std::make_pair(H, W)];   // a training loop invalidates the table every step: no new
            if (!buf) {                                          // allocation per update
                void* p = nullptr;
                DAD_TRY(alloc(&p, rows * D() * 4));
                buf = reinterpret_cast<float*>(p);
            }
            tab = buf;
            DAD_TRY(pos_table(P("pretrained.pos_embed"), P("pretrained.cls_token"), P("pretrained.patch_embed.proj.bias"),
                              tab, D(), H, W, st));
        }
        return DAD_OK;
    }

    int capture(const char* name, const void* src, bool src_bf16, long long numel, cudaStream_t st) {
        auto it = captures.find(name);
        if (it == captures.end() || !src) return DAD_OK;
        DAD_REQUIRE(it->second.second == numel, "debug capture %s: buffer has %lld elements, tensor has %lld", name,
                    it->second.second, numel);
        if (src_bf16) {
            to_f32_kernel<<<static_cast<unsigned>(cdivl(numel, 256)), 256, 0, st>>>(reinterpret_cast<const bf16*>(src),
                                                                                   it->second.first, numel);
            DAD_CHECK_LAUNCH();
        } else {
            DAD_CHECK_CUDA(cudaMemcpyAsync(it->second.first, src, numel * 4, cudaMemcpyDeviceToDevice, st));
        }
        return DAD_OK;
    }

    int gemm(int mode, const GemmProblem& p, bool dry, cudaStream_t st) {
        if (dry) return DAD_OK;
        {
            char lab[128];
            snprintf(lab, sizeof(lab), "gemm conv=%d M=%d K=%d N=%d Kp=%d B=%d H=%d W=%d C=%d taps=%d scat=%d", p.conv, p.M,
                     p.K, p.N, p.Kp, p.B, p.H, p.W, p.C, p.taps, p.epi.scat_k);
            debug_label(lab);
        }
        return mode == 0 ? gemm_tc(p, st) : gemm_simt(p, st);
    }

    // conv3x3 / conv1x1 (stride 1) on an NHWC tensor through the GEMM engine
    int conv(int mode, const void* in, int B, int H, int W, int C, const Mat& m, int taps, const Epilogue& e, bool dry,
             cudaStream_t st, int stride = 1) {
        GemmProblem p;
        p.A = in; p.conv = 1; p.B = B; p.H = H; p.W = W; p.C = C; p.taps = taps; p.ldp = C; p.stride = stride;
        p.Wt = m.w[mode]; p.N = m.N; p.Kp = m.Kp;
        p.epi = e;
        if (p.epi.ldc == 0) p.epi.ldc = m.N;
        return gemm(mode, p, dry, st);
    }
    int linear(int mode, const void* in, long long M, int K, const Mat& m, const Epilogue& e, bool dry, cudaStream_t st) {
        GemmProblem p;
        p.A = in; p.conv = 0; p.M = static_cast<int>(M); p.K = K; p.lda = K;
        p.Wt = m.w[mode]; p.N = m.N; p.Kp = m.Kp;
        p.epi = e;
        if (p.epi.ldc == 0) p.epi.ldc = m.N;
        return gemm(mode, p, dry, st);
    }

    // one FeatureFusionBlock (util/blocks.py:129-146); x0 = upsampled path (may be null), x1 = lateral
    // `lat` / `lat_relu`: the tensor entering the first RCU and its ReLU copy.
    int fusion(int mode, int r, bool has_path, const void* path, const void* lat, const void* lat_relu, int B, int H, int W, int Ho,
               int Wo, void** out, Arena& ar, bool dry, cudaStream_t st) {
        const int F = desc.features;
        const size_t es = mode == 0 ? 2 : 4;
        const int bf = mode == 0;
        const size_t n = static_cast<size_t>(B) * H * W * F;
        const std::string q = "depth_head.scratch.refinenet" + std::to_string(r + 1) + ".";
        // all scratch of this block up front (identical in the dry run), then launch
        const size_t no = static_cast<size_t>(B) * Ho * Wo * F;
        void* t1 = ar.take(n * es);
        void* s = has_path ? ar.take(n * es) : nullptr;
        void* sr = has_path ? ar.take(n * es) : nullptr;
        void* u = ar.take(n * es);
        void* res = ar.take(no * es);
        void* tmp = ar.take((mode == 0 ? n : no) * es);  // bf16: 1x1 conv output at low res; fp32: upsampled map
        DAD_REQUIRE(!ar.overflow, "forward: workspace arena overflow in fusion block %d", r + 1);
        const void* sum = lat;
        const void* sum_relu = lat_relu;
        if (has_path) {  // output = path + RCU1(lat)
            Epilogue e1; e1.bias = P(q + "resConfUnit1.conv1.bias"); e1.act = ACT_RELU; e1.out = t1; e1.out_bf16 = bf;
            DAD_TRY(conv(mode, lat_relu, B, H, W, F, rcu[r][0][0], 9, e1, dry, st));
            Epilogue e2; e2.bias = P(q + "resConfUnit1.conv2.bias"); e2.res1 = lat; e2.res1_bf16 = bf; e2.res2 = path;
            e2.res2_bf16 = bf; e2.out = s; e2.out_bf16 = bf; e2.out_relu = sr;
            DAD_TRY(conv(mode, t1, B, H, W, F, rcu[r][0][1], 9, e2, dry, st));
            sum = s; sum_relu = sr;
        }
        {
            Epilogue e1; e1.bias = P(q + "resConfUnit2.conv1.bias"); e1.act = ACT_RELU; e1.out = t1; e1.out_bf16 = bf;
            DAD_TRY(conv(mode, sum_relu, B, H, W, F, rcu[r][1][0], 9, e1, dry, st));
            Epilogue e2; e2.bias = P(q + "resConfUnit2.conv2.bias"); e2.res1 = sum; e2.res1_bf16 = bf; e2.out = u;
            e2.out_bf16 = bf;
            DAD_TRY(conv(mode, t1, B, H, W, F, rcu[r][1][1], 9, e2, dry, st));
        }
        Epilogue eo; eo.bias = P(q + "out_conv.bias"); eo.out_bf16 = bf;
        if (mode == 0) {
            // 1x1 conv commutes with bilinear resampling (both linear; weights sum to 1): conv at the low
            // resolution (4x fewer FLOPs), then resample.  fp32 mode keeps the reference order.
            eo.out = tmp;
            DAD_TRY(conv(mode, u, B, H, W, F, out_conv[r], 1, eo, dry, st));
            debug_label("fusion bilinear (bf16 order)");
            if (!dry) DAD_TRY(bilinear_nhwc(tmp, res, bf, B, H, W, Ho, Wo, F, st));
        } else {
            debug_label("fusion bilinear (reference order)");
            if (!dry) DAD_TRY(bilinear_nhwc(u, tmp, bf, B, H, W, Ho, Wo, F, st));
            eo.out = res;
            DAD_TRY(conv(mode, tmp, B, Ho, Wo, F, out_conv[r], 1, eo, dry, st));
        }
        *out = res;
        return DAD_OK;
    }

    int forward(const float* x, int B, int H, int W, int mode, float* depth_out, float* feat_out, void* ws,
                size_t ws_bytes, size_t* ws_needed, cudaStream_t st) {
        const bool dry = ws_needed != nullptr;
        DAD_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (bf16) or 1 (fp32)");
        if (!dry) {  // size check BEFORE anything is launched
            size_t need = 0;
            DAD_TRY(forward(nullptr, B, H, W, mode, nullptr, nullptr, nullptr, 0, &need, nullptr));
            if (ws_bytes < need)
                return set_error(DAD_ERR_WORKSPACE, "forward workspace too small: need %zu bytes, got %zu", need, ws_bytes);
        }
        DAD_REQUIRE(B > 0 && H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0,
                    "input must be [B,3,H,W] with H, W positive multiples of 14 (got B=%d H=%d W=%d)", B, H, W);
        if (!dry) {
            DAD_REQUIRE(x && depth_out, "forward: null input/output");
            const auto pit = pos_tables.find(std::make_pair(H, W));
            const bool ready = packed[mode] && pit != pos_tables.end() && pit->second != nullptr;
            DAD_REQUIRE(ready, "forward: call dad_model_prepare(mode, H, W) first");
            DAD_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "workspace must be 1024-byte aligned");
        }
        const int Dm = D(), L = desc.depth, F = desc.features, heads = desc.num_heads;
        const int* oc = desc.out_channels;
        const int ph = H / 14, pw = W / 14, np = ph * pw, T = np + 1;
        const long long M = static_cast<long long>(B) * T, Mp = static_cast<long long>(B) * np;
        DAD_REQUIRE(M < (1LL << 31) / 4, "batch too large for 32-bit row indices");
        const int bf = mode == 0;
        const size_t es = bf ? 2 : 4;
        Arena ar(ws, ws_bytes, dry);
        const std::string p = "pretrained.";

        // ---- encoder -------------------------------------------------------------------------
        float* xres = reinterpret_cast<float*>(ar.take(M * Dm * 4));
        void* ape = ar.take(M * PATCH_KP * es);
        void* ln = ar.take(M * Dm * es);
        void* qkvb = ar.take(M * 3 * Dm * es);
        void* att = ar.take(M * Dm * es);
        // optional parts (read from the parameter set, so the workspace query works before dad_model_prepare as well)
        const bool readout = has_readout(), swiglu_ffn = swiglu_hidden() > 0;
        const int Hd = swiglu_hidden();
        void* hid = ar.take(M * (swiglu_ffn ? 2 * Hd : 4 * Dm) * es);
        void* gated = swiglu_ffn ? ar.take(M * Hd * es) : nullptr;
        void* tapbuf[4];
        for (int j = 0; j < 4; ++j) tapbuf[j] = ar.take(Mp * Dm * es);
        void* tapraw = readout ? ar.take(Mp * Dm * es) : nullptr;       // normalised patch tokens before the readout
        void* clsbuf = readout ? ar.take(static_cast<size_t>(B) * Dm * es) : nullptr;
        void* catbuf = readout ? ar.take(Mp * 2 * Dm * es) : nullptr;   // [patch | cls] rows
        DAD_REQUIRE(!ar.overflow, "forward: workspace arena overflow (encoder)");
        if (!dry) {
            debug_label("patch_im2col");
            DAD_TRY(patch_im2col(x, ape, bf, B, H, W, PATCH_KP, st));
            Epilogue e; e.rowtab = pos_tables[std::make_pair(H, W)]; e.rowtab_period = T; e.out = xres;
            DAD_TRY(linear(mode, ape, M, PATCH_KP, patch, e, dry, st));
            DAD_TRY(capture("tokens", xres, false, M * Dm, st));
            int tj = 0;
            for (int i = 0; i < L; ++i) {
                const std::string b = p + "blocks." + std::to_string(i) + ".";
                debug_label(("block " + std::to_string(i) + " ln/attn").c_str());
                DAD_TRY(layernorm(xres, P(b + "norm1.weight"), P(b + "norm1.bias"), ln, bf, nullptr, M, Dm, 1, 1, 0, LN_EPS, st));
                Epilogue eq; eq.bias = bqkv_scaled + static_cast<long long>(i) * 3 * Dm; eq.out = qkvb; eq.out_bf16 = bf;
                DAD_TRY(linear(mode, ln, M, Dm, qkv[i], eq, dry, st));
                debug_label(("block " + std::to_string(i) + " attention").c_str());
                DAD_TRY(attention(qkvb, att, bf, B, T, heads, st));
                Epilogue ep; ep.bias = P(b + "attn.proj.bias"); ep.gamma = P(b + "ls1.gamma"); ep.res1 = xres; ep.out = xres;
                DAD_TRY(linear(mode, att, M, Dm, proj[i], ep, dry, st));
                debug_label(("block " + std::to_string(i) + " ln2").c_str());
                DAD_TRY(layernorm(xres, P(b + "norm2.weight"), P(b + "norm2.bias"), ln, bf, nullptr, M, Dm, 1, 1, 0, LN_EPS, st));
                if (swiglu_ffn) {
                    Epilogue e1; e1.bias = P(b + "mlp.w12.bias"); e1.out = hid; e1.out_bf16 = bf;
                    DAD_TRY(linear(mode, ln, M, Dm, w12[i], e1, dry, st));
                    debug_label("swiglu gate");
                    DAD_TRY(swiglu(hid, gated, bf, M, Hd, st));
                    Epilogue e2; e2.bias = P(b + "mlp.w3.bias"); e2.gamma = P(b + "ls2.gamma"); e2.res1 = xres; e2.out = xres;
                    DAD_TRY(linear(mode, gated, M, Hd, w3[i], e2, dry, st));
                } else {
                    Epilogue e1; e1.bias = P(b + "mlp.fc1.bias"); e1.act = ACT_GELU; e1.out = hid; e1.out_bf16 = bf;
                    DAD_TRY(linear(mode, ln, M, Dm, fc1[i], e1, dry, st));
                    Epilogue e2; e2.bias = P(b + "mlp.fc2.bias"); e2.gamma = P(b + "ls2.gamma"); e2.res1 = xres; e2.out = xres;
                    DAD_TRY(linear(mode, hid, M, 4 * Dm, fc2[i], e2, dry, st));
                }
                if (i == 0) DAD_TRY(capture("block0", xres, false, M * Dm, st));
                if (i == L - 1) DAD_TRY(capture("block_last", xres, false, M * Dm, st));
                if (tj < 4 && i == desc.taps[tj]) {
                    // final LayerNorm on the tapped residual, cls row dropped (dinov2.py:310-312)
                    debug_label("tap layernorm");
                    DAD_TRY(layernorm(xres, P(p + "norm.weight"), P(p + "norm.bias"), readout ? tapraw : tapbuf[tj], bf,
                                      (tj == 3) ? feat_out : nullptr, Mp, Dm, np, T, 1, LN_EPS, st));
                    if (readout) {
                        // dpt.py:153-156: x = GELU(Linear(cat(x, cls.expand_as(x)))); features[3][0] (feat_out) stays the
                        // plain normalised patch tokens
                        const std::string r = "depth_head.readout_projects." + std::to_string(tj) + ".0.";
                        debug_label("readout: cls layernorm + concat + project");
                        DAD_TRY(layernorm(xres, P(p + "norm.weight"), P(p + "norm.bias"), clsbuf, bf, nullptr, B, Dm, 1, T, 0,
                                          LN_EPS, st));
                        DAD_TRY(concat_cls(tapraw, clsbuf, catbuf, bf, B, np, Dm, st));
                        Epilogue er; er.bias = P(r + "bias"); er.act = ACT_GELU; er.out = tapbuf[tj]; er.out_bf16 = bf;
                        DAD_TRY(linear(mode, catbuf, Mp, 2 * Dm, readout_proj[tj], er, dry, st));
                    }
                    ++tj;
                }
            }
            DAD_REQUIRE(tj == 4, "taps must be increasing block indices < depth");
        }

        // ---- DPT head ------------------------------------------------------------------------
        const std::string h = "depth_head.", s = h + "scratch.";
        const int hs[4] = {4 * ph, 2 * ph, ph, (ph + 2 - 3) / 2 + 1};
        const int wsz[4] = {4 * pw, 2 * pw, pw, (pw + 2 - 3) / 2 + 1};
        void* lrn[4];
        void* lrn_relu[4];
        for (int j = 0; j < 4; ++j) {
            void* pj = ar.take(Mp * oc[j] * es);
            Epilogue e; e.bias = P(h + "projects." + std::to_string(j) + ".bias"); e.out = pj; e.out_bf16 = bf;
            DAD_TRY(linear(mode, tapbuf[j], Mp, Dm, projects[j], e, dry, st));
            void* rj = pj;
            if (j == 0 || j == 1) {
                const int k = j == 0 ? 4 : 2;
                rj = ar.take(static_cast<size_t>(B) * hs[j] * wsz[j] * oc[j] * es);
                Epilogue es_; es_.bias = P(h + "resize_layers." + std::to_string(j) + ".bias"); es_.out = rj; es_.out_bf16 = bf;
                es_.ldc = oc[j]; es_.scat_k = k; es_.scat_CoP = j == 0 ? CoP0 : CoP1; es_.scat_Co = oc[j];
                es_.scat_H = ph; es_.scat_W = pw;
                if (mode == 0 && oc[j] % 64 == 0)  // 1x1 "conv" + 5-D TMA scatter store (gemm_tc.cu)
                    DAD_TRY(conv(mode, pj, B, ph, pw, oc[j], j == 0 ? resize0 : resize1, 1, es_, dry, st));
                else
                    DAD_TRY(linear(mode, pj, Mp, oc[j], j == 0 ? resize0 : resize1, es_, dry, st));
            } else if (j == 3) {
                const int Cp = cdiv(oc[3], 64) * 64;
                const long long rows = static_cast<long long>(B) * hs[3] * wsz[3];
                Epilogue e3; e3.bias = P(h + "resize_layers.3.bias"); e3.out_bf16 = bf;
                if (mode == 0) {
                    // implicit GEMM straight from the NHWC map: the TMA box walks the input with element stride 2
                    rj = ar.take(rows * oc[3] * es);
                    e3.out = rj;
                    DAD_TRY(conv(mode, pj, B, ph, pw, oc[3], resize3, 9, e3, dry, st, 2));
                } else {
                    void* col = ar.take(rows * 9 * Cp * es);
                    rj = ar.take(rows * oc[3] * es);
                    e3.out = rj;
                    debug_label("im2col_s2");
                    if (!dry) DAD_TRY(im2col_s2(pj, col, bf, B, ph, pw, oc[3], Cp, st));
                    DAD_TRY(linear(mode, col, rows, 9 * Cp, resize3, e3, dry, st));
                }
            }
            const size_t n = static_cast<size_t>(B) * hs[j] * wsz[j] * F;
            lrn[j] = ar.take(n * es);
            lrn_relu[j] = ar.take(n * es);
            DAD_REQUIRE(!ar.overflow, "forward: workspace arena overflow (reassemble %d)", j);
            Epilogue er; er.out = lrn[j]; er.out_bf16 = bf; er.out_relu = lrn_relu[j];
            DAD_TRY(conv(mode, rj, B, hs[j], wsz[j], oc[j], layer_rn[j], 9, er, dry, st));
            if (!dry) DAD_TRY(capture(("layer_rn" + std::to_string(j + 1)).c_str(), lrn[j], bf, n, st));
        }
        void *p4, *p3, *p2, *p1;
        DAD_TRY(fusion(mode, 3, false, nullptr, lrn[3], lrn_relu[3], B, hs[3], wsz[3], hs[2], wsz[2], &p4, ar, dry, st));
        DAD_TRY(fusion(mode, 2, true, p4, lrn[2], lrn_relu[2], B, hs[2], wsz[2], hs[1], wsz[1], &p3, ar, dry, st));
        DAD_TRY(fusion(mode, 1, true, p3, lrn[1], lrn_relu[1], B, hs[1], wsz[1], hs[0], wsz[0], &p2, ar, dry, st));
        DAD_TRY(fusion(mode, 0, true, p2, lrn[0], lrn_relu[0], B, hs[0], wsz[0], 2 * hs[0], 2 * wsz[0], &p1, ar, dry, st));
        const int H1 = 2 * hs[0], W1 = 2 * wsz[0];
        if (!dry) {
            DAD_TRY(capture("path_4", p4, bf, static_cast<long long>(B) * hs[2] * wsz[2] * F, st));
            DAD_TRY(capture("path_1", p1, bf, static_cast<long long>(B) * H1 * W1 * F, st));
        }
        const int F2 = F / 2;
        void* o1 = ar.take(static_cast<size_t>(B) * H1 * W1 * F2 * es);
        Epilogue eo1; eo1.bias = P(s + "output_conv1.bias"); eo1.out = o1; eo1.out_bf16 = bf;
        DAD_TRY(conv(mode, p1, B, H1, W1, F, output_conv1, 9, eo1, dry, st));
        // K16 + K17.  Measured and rejected in round 2: (a) running upsample + head conv per L2-sized image
        // group through one small buffer keeps the 2.2 GB upsampled tensor out of HBM but is no faster (the head conv is
        // bound by shared-memory operand bandwidth at N = 32, the per-image launches lose tail efficiency); (b) gathering
        // the bilinear taps inside the conv's A-operand load adds shared-memory traffic to that same bottleneck.
        void* up = ar.take(static_cast<size_t>(B) * H * W * F2 * es);
        if (!dry && getenv("DAD_DEBUG_SYNC"))
            fprintf(stderr, "dad[debug]: ws=%p bytes=%zu used=%zu o1=%p up=%p (H1=%d W1=%d H=%d W=%d F2=%d es=%zu)\n", ws,
                    ws_bytes, ar.used, o1, up, H1, W1, H, W, F2, es);
        float* t32 = mode == 1 ? reinterpret_cast<float*>(ar.take(static_cast<size_t>(B) * H * W * 32 * 4)) : nullptr;
        DAD_REQUIRE(!ar.overflow, "forward: workspace arena overflow (head)");
        debug_label("head bilinear");
        if (!dry) DAD_TRY(bilinear_nhwc(o1, up, bf, B, H1, W1, H, W, F2, st));
        if (mode == 0) {
            // conv3x3 -> ReLU -> conv1x1 -> ReLU (-> F.relu) fused in the GEMM epilogue
            Epilogue eh; eh.bias = P(s + "output_conv2.0.bias"); eh.head_w = P(s + "output_conv2.2.weight");
            eh.head_b = P(s + "output_conv2.2.bias"); eh.head_out = depth_out;
            DAD_TRY(conv(mode, up, B, H, W, F2, output_conv2_0, 9, eh, dry, st));
        } else {
            Epilogue eh; eh.bias = P(s + "output_conv2.0.bias"); eh.act = ACT_RELU; eh.out = t32;
            DAD_TRY(conv(mode, up, B, H, W, F2, output_conv2_0, 9, eh, dry, st));
            if (!dry) DAD_TRY(head1x1(t32, P(s + "output_conv2.2.weight"), P(s + "output_conv2.2.bias"), depth_out,
                                      static_cast<long long>(B) * H * W, st));
        }
        if (dry) *ws_needed = ar.used + 1024;
        return DAD_OK;
    }
};

#include "train.inl"

namespace {

int train_checks(Model& m, int B, int H, int W, int mode) {
    DAD_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (bf16) or 1 (fp32)");
    DAD_REQUIRE(B > 0 && H > 0 && W > 0 && H % 14 == 0 && W % 14 == 0,
                "input must be [B,3,H,W] with H, W positive multiples of 14 (got B=%d H=%d W=%d)", B, H, W);
    DAD_REQUIRE(static_cast<long long>(B) * (1 + (H / 14) * (W / 14)) < (1LL << 31) / 4, "batch too large for 32-bit row indices");
    (void)m;
    return DAD_OK;
}

int train_workspace(Model& m, int B, int H, int W, int mode, size_t* need) {
    DAD_TRY(train_checks(m, B, H, W, mode));
    size_t peak = 0;
    {
        Bump ar(nullptr, 0, true);
        Tape t;
        Trainer tr(m, B, H, W, mode, true, nullptr);
        DAD_TRY(tr.forward(nullptr, nullptr, nullptr, ar, t));
        peak = ar.peak;
    }
    {
        Bump ar(nullptr, 0, true);
        Tape t;
        Trainer tr(m, B, H, W, mode, true, nullptr);
        DAD_TRY(tr.backward(nullptr, nullptr, ar, t));
        if (ar.peak > peak) peak = ar.peak;
    }
    *need = peak + 1024;
    return DAD_OK;
}

int train_ready(Model& m, int B, int H, int W, int mode, void* ws, size_t ws_bytes) {
    size_t need = 0;
    DAD_TRY(train_workspace(m, B, H, W, mode, &need));
    if (ws_bytes < need) return set_error(DAD_ERR_WORKSPACE, "training workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    const auto pit = m.pos_tables.find(std::make_pair(H, W));
    DAD_REQUIRE(m.packed[1] && m.packed[mode] && pit != m.pos_tables.end() && pit->second != nullptr,
                "call dad_model_prepare(mode, H, W) first (training in mode 0 also needs the fp32 pack: prepare both modes)");
    DAD_REQUIRE(ws && (reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "workspace must be 1024-byte aligned");
    return DAD_OK;
}

}  // namespace

}  // namespace dad

// ====================================================================== C ABI (model part)
#include "../../include/dad_b200.h"

extern "C" {

int dad_model_create(const dad_model_desc* d, dad_model** out) {
    if (!d || !out) return dad::set_error(DAD_ERR_INVALID, "dad_model_create: null argument");
    if (d->embed_dim <= 0 || d->depth <= 0 || d->num_heads <= 0 || d->embed_dim != d->num_heads * 64)
        return dad::set_error(DAD_ERR_UNSUPPORTED, "dad_model_create: embed_dim=%d heads=%d (head_dim must be 64)",
                              d->embed_dim, d->num_heads);
    dad::ModelDesc m{};
    m.embed_dim = d->embed_dim; m.depth = d->depth; m.num_heads = d->num_heads; m.features = d->features;
    for (int i = 0; i < 4; ++i) { m.taps[i] = d->taps[i]; m.out_channels[i] = d->out_channels[i]; }
    *out = reinterpret_cast<dad_model*>(new dad::Model(m));
    return DAD_OK;
}

void dad_model_destroy(dad_model* m) { delete reinterpret_cast<dad::Model*>(m); }

int dad_model_set_weight(dad_model* m, const char* name, const float* dev_ptr, int64_t numel, void* stream) {
    if (!m) return dad::set_error(DAD_ERR_INVALID, "null model");
    return reinterpret_cast<dad::Model*>(m)->set_weight(name, dev_ptr, numel, reinterpret_cast<cudaStream_t>(stream));
}

int dad_model_prepare(dad_model* m, int mode, int H, int W, void* stream) {
    if (!m) return dad::set_error(DAD_ERR_INVALID, "null model");
    return reinterpret_cast<dad::Model*>(m)->prepare(mode, H, W, reinterpret_cast<cudaStream_t>(stream));
}

int dad_model_debug_capture(dad_model* m, const char* name, float* dst, int64_t numel) {
    if (!m || !name) return dad::set_error(DAD_ERR_INVALID, "null argument");
    auto* mm = reinterpret_cast<dad::Model*>(m);
    if (!dst) mm->captures.erase(name);
    else mm->captures[name] = {dst, numel};
    return DAD_OK;
}

int dad_model_set_grad(dad_model* m, const char* name, float* dev_ptr, int64_t numel) {
    if (!m || !name) return dad::set_error(DAD_ERR_INVALID, "null argument");
    auto* mm = reinterpret_cast<dad::Model*>(m);
    if (!dev_ptr) { mm->grads.erase(name); return DAD_OK; }
    auto it = mm->master.find(name);
    if (it == mm->master.end()) return dad::set_error(DAD_ERR_INVALID, "set_grad: unknown parameter %s", name);
    if (it->second.second != numel)
        return dad::set_error(DAD_ERR_INVALID, "set_grad: %s has %lld elements, gradient buffer has %lld", name,
                              it->second.second, static_cast<long long>(numel));
    mm->grads[name] = {dev_ptr, numel};
    return DAD_OK;
}

size_t dad_train_workspace_bytes(dad_model* m, int B, int H, int W, int mode) {
    if (!m) return 0;
    size_t need = 0;
    if (dad::train_workspace(*reinterpret_cast<dad::Model*>(m), B, H, W, mode, &need) != DAD_OK) return 0;
    return need;
}

int dad_forward_train(dad_model* m, const float* x, int B, int H, int W, int mode, float* depth_out, float* feat_out,
                      void* workspace, size_t workspace_bytes, void* stream) {
    if (!m) return dad::set_error(DAD_ERR_INVALID, "null model");
    if (!x || !depth_out) return dad::set_error(DAD_ERR_INVALID, "forward_train: null input/output");
    auto& mm = *reinterpret_cast<dad::Model*>(m);
    DAD_TRY(dad::train_ready(mm, B, H, W, mode, workspace, workspace_bytes));
    dad::Bump ar(workspace, workspace_bytes, false);
    dad::Tape t;
    dad::Trainer tr(mm, B, H, W, mode, false, reinterpret_cast<cudaStream_t>(stream));
    return tr.forward(x, depth_out, feat_out, ar, t);
}

int dad_backward(dad_model* m, int B, int H, int W, int mode, const float* grad_depth, const float* grad_feat, void* workspace,
                 size_t workspace_bytes, void* stream) {
    if (!m) return dad::set_error(DAD_ERR_INVALID, "null model");
    auto& mm = *reinterpret_cast<dad::Model*>(m);
    DAD_TRY(dad::train_ready(mm, B, H, W, mode, workspace, workspace_bytes));
    dad::Bump ar(workspace, workspace_bytes, false);
    dad::Tape t;
    dad::Trainer tr(mm, B, H, W, mode, false, reinterpret_cast<cudaStream_t>(stream));
    return tr.backward(grad_depth, grad_feat, ar, t);
}

size_t dad_forward_workspace_bytes(dad_model* m, int B, int H, int W, int mode) {
    if (!m) return 0;
    size_t need = 0;
    if (reinterpret_cast<dad::Model*>(m)->forward(nullptr, B, H, W, mode, nullptr, nullptr, nullptr, 0, &need, nullptr) != DAD_OK)
        return 0;
    return need;
}

int dad_forward(dad_model* m, const float* x, int B, int H, int W, int mode, float* depth_out, float* feat_out,
                void* workspace, size_t workspace_bytes, void* stream) {
    if (!m) return dad::set_error(DAD_ERR_INVALID, "null model");
    return reinterpret_cast<dad::Model*>(m)->forward(x, B, H, W, mode, depth_out, feat_out, workspace, workspace_bytes,
                                                     nullptr, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
