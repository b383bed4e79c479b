// Training forward (activations kept on a tape) and backward of the DepthAnythingV2 student (reference: autograd over dpt.py:150-225, dinov2.py:212-321,
// util/blocks.py:29-148, driven by tools/train_distillation.py:1556-1575).  Included by model.cu inside namespace dad.
//
// One caller-owned workspace holds the tape (written by forward_train, read by backward) followed by the backward
// scratch; both calls lay it out with the same deterministic bump sequence, so no pointers are carried between them.
// Parameter gradients ACCUMULATE into the fp32 buffers registered with dad_model_set_grad (unregistered = frozen).
//
// mode 1 (fp32): every tensor fp32, every contraction on the FFMA engine (gemm_simt / sgemm) - the verification path.
// mode 0 (bf16): activations and activation gradients bf16 (residual stream, its gradient and all parameter gradients
//   fp32); forward GEMMs / convs / attention on the tcgen05 engine; backward: linear and conv data gradients as gemm_tc /
//   conv_tc2 launches with transposed / flipped bf16 weights, linear and conv weight gradients as split-K gemm_tc launches
//   over K-major transposes (dY^T, X^T / im2col^T) that reduce-add fp32 partial tiles through TMA.  The attention
//   backward and the small reassemble stage (ConvTranspose, stride-2 conv) still run on the fp32 engine via conversions.

namespace {

struct Bump {  // bump allocator with release-to-mark and a high-water mark
    uint8_t* base;
    size_t cap;
    bool dry;
    size_t used = 0, peak = 0;
    bool overflow = false;
    Bump(void* b, size_t c, bool d) : base(reinterpret_cast<uint8_t*>(b)), cap(c), dry(d) {}
    void* bytes(size_t nbytes) {
        used = (used + 1023) & ~size_t(1023);
        void* p = dry ? nullptr : base + used;
        used += nbytes;
        if (used > peak) peak = used;
        if (!dry && used > cap) { overflow = true; p = nullptr; }
        return p;
    }
    float* f(size_t n) { return reinterpret_cast<float*>(bytes(n * 4)); }
};

// x0 / x1 / xfinal / depth are fp32 in both modes; every other tape tensor has the mode's activation type
struct BlockTape { float *x0, *x1; void *n1, *qkv, *att, *y1, *n2, *hpre, *h, *y2; };
struct FusionTape {
    void *t1a = nullptr, *s = nullptr, *sr = nullptr, *t1b = nullptr, *u = nullptr, *tmp = nullptr, *res = nullptr;
    const void *lat = nullptr, *lat_relu = nullptr, *path = nullptr;
    int H = 0, W = 0, Ho = 0, Wo = 0;
    bool has_path = false;
};
struct Tape {
    void* ape = nullptr;
    std::vector<BlockTape> blk;
    float* xfinal = nullptr;
    void *tap[4], *pj[4], *rj[4], *lrn[4], *lrn_relu[4];
    // use_clstoken readout (dpt.py:153-156): normalised patch tokens, normalised class tokens, [patch | cls] rows, pre-GELU
    void *tapraw[4] = {nullptr, nullptr, nullptr, nullptr}, *cls[4] = {nullptr, nullptr, nullptr, nullptr};
    void *cat[4] = {nullptr, nullptr, nullptr, nullptr}, *rpre[4] = {nullptr, nullptr, nullptr, nullptr};
    FusionTape fu[4];
    void *o1 = nullptr, *up = nullptr, *t32 = nullptr;
    float* depth = nullptr;
    int hs[4], wsz[4];
};

}  // namespace

struct Trainer {
    Model& m;
    int B, H, W, mode, bf;
    size_t es;
    bool dry;
    cudaStream_t st;
    int Dm, L, F, heads, ph, pw, np, T;
    long long M, Mp;
    const int* oc;
    int Hd = 0;             // SwiGLU hidden width (ViT-g); 0 = Mlp with GELU
    bool readout = false;   // use_clstoken readout projections present

    Trainer(Model& model, int B_, int H_, int W_, int mode_, bool dry_, cudaStream_t st_)
        : m(model), B(B_), H(H_), W(W_), mode(mode_), bf(mode_ == 0), es(mode_ == 0 ? 2 : 4), dry(dry_), st(st_) {
        Dm = m.D(); L = m.desc.depth; F = m.desc.features; heads = m.desc.num_heads; oc = m.desc.out_channels;
        ph = H / 14; pw = W / 14; np = ph * pw; T = np + 1;
        M = static_cast<long long>(B) * T; Mp = static_cast<long long>(B) * np;
        Hd = m.swiglu_hidden(); readout = m.has_readout();
    }

#define RUN(expr) do { if (!dry) DAD_TRY(expr); } while (0)

    void* a(Bump& ar, size_t n) const { return ar.bytes(n * es); }   // n elements of the activation type
    const float* zeros = nullptr;   // [8192] 0 / 1 vectors for the TMA epilogues of the backward GEMMs (mode 0)
    const float* ones = nullptr;
    const float* eighths = nullptr;

    float* G(const std::string& name) const {
        auto it = m.grads.find(name);
        return it == m.grads.end() ? nullptr : it->second.first;
    }

    void plan(Bump& ar, Tape& t) const {
        t.ape = a(ar, M * PATCH_KP);
        t.blk.resize(L);
        for (int i = 0; i < L; ++i) {
            BlockTape& b = t.blk[i];
            b.x0 = ar.f(M * Dm); b.n1 = a(ar, M * Dm); b.qkv = a(ar, M * 3 * Dm); b.att = a(ar, M * Dm); b.y1 = a(ar, M * Dm);
            // Mlp: fc1 pre-activation and GELU output [M, 4D]; SwiGLU: x12 = w12(n2) [M, 2*Hd] and the gated product [M, Hd]
            b.x1 = ar.f(M * Dm); b.n2 = a(ar, M * Dm); b.hpre = a(ar, M * (Hd ? 2 * Hd : 4 * Dm)); b.h = a(ar, M * (Hd ? Hd : 4 * Dm));
            b.y2 = a(ar, M * Dm);
        }
        t.xfinal = ar.f(M * Dm);
        const int hs[4] = {4 * ph, 2 * ph, ph, (ph + 2 - 3) / 2 + 1};
        const int wz[4] = {4 * pw, 2 * pw, pw, (pw + 2 - 3) / 2 + 1};
        for (int j = 0; j < 4; ++j) {
            t.hs[j] = hs[j]; t.wsz[j] = wz[j];
            t.tap[j] = a(ar, Mp * Dm);
            if (readout) {
                t.tapraw[j] = a(ar, Mp * Dm); t.cls[j] = a(ar, static_cast<size_t>(B) * Dm); t.cat[j] = a(ar, Mp * 2 * Dm);
                t.rpre[j] = a(ar, Mp * Dm);
            }
            t.pj[j] = a(ar, Mp * oc[j]);
            t.rj[j] = (j == 2) ? t.pj[j] : a(ar, static_cast<size_t>(B) * hs[j] * wz[j] * oc[j]);
            const size_t n = static_cast<size_t>(B) * hs[j] * wz[j] * F;
            t.lrn[j] = a(ar, n);
            t.lrn_relu[j] = a(ar, n);
        }
        for (int r = 3; r >= 0; --r) {
            FusionTape& f = t.fu[r];
            f.H = hs[r]; f.W = wz[r];
            f.Ho = r > 0 ? hs[r - 1] : 2 * hs[0];
            f.Wo = r > 0 ? wz[r - 1] : 2 * wz[0];
            f.has_path = r != 3;
            const size_t n = static_cast<size_t>(B) * f.H * f.W * F, no = static_cast<size_t>(B) * f.Ho * f.Wo * F;
            f.lat = t.lrn[r]; f.lat_relu = t.lrn_relu[r];
            f.path = f.has_path ? t.fu[r + 1].res : nullptr;
            if (f.has_path) { f.t1a = a(ar, n); f.s = a(ar, n); f.sr = a(ar, n); }
            else { f.s = t.lrn[r]; f.sr = t.lrn_relu[r]; }
            f.t1b = a(ar, n); f.u = a(ar, n); f.tmp = a(ar, no); f.res = a(ar, no);
        }
        const int H1 = 2 * hs[0], W1 = 2 * wz[0], F2 = F / 2;
        t.o1 = a(ar, static_cast<size_t>(B) * H1 * W1 * F2);
        t.up = a(ar, static_cast<size_t>(B) * H * W * F2);
        t.t32 = a(ar, static_cast<size_t>(B) * H * W * 32);
        t.depth = ar.f(static_cast<size_t>(B) * H * W);
    }

    // ------------------------------------------------------------------------------------ forward
    Epilogue epi(const float* bias, void* out) const {
        Epilogue e; e.bias = bias; e.out = out; e.out_bf16 = bf;
        return e;
    }

    int forward(const float* x, float* depth_out, float* feat_out, Bump& ar, Tape& t) {
        plan(ar, t);
        // transient im2col buffer of the stride-2 reassemble conv, fp32 engine only (lies in the backward-scratch region)
        const int Cp3 = cdiv(oc[3], 64) * 64;
        const long long rows3 = static_cast<long long>(B) * t.hs[3] * t.wsz[3];
        float* col = mode == 1 ? ar.f(rows3 * 9 * Cp3) : nullptr;
        if (dry) return DAD_OK;
        DAD_REQUIRE(!ar.overflow, "forward_train: workspace too small for the activation tape");
        const std::string p = "pretrained.";
        DAD_TRY(patch_im2col(x, t.ape, bf, B, H, W, PATCH_KP, st));
        {
            Epilogue e; e.rowtab = m.pos_tables[std::make_pair(H, W)]; e.rowtab_period = T; e.out = t.blk[0].x0;
            DAD_TRY(m.linear(mode, t.ape, M, PATCH_KP, m.patch, e, false, st));
        }
        int tj = 0;
        for (int i = 0; i < L; ++i) {
            const std::string b = p + "blocks." + std::to_string(i) + ".";
            BlockTape& bt = t.blk[i];
            float* xnext = (i + 1 < L) ? t.blk[i + 1].x0 : t.xfinal;
            DAD_TRY(layernorm(bt.x0, m.P(b + "norm1.weight"), m.P(b + "norm1.bias"), bt.n1, bf, nullptr, M, Dm, 1, 1, 0, LN_EPS, st));
            DAD_TRY(m.linear(mode, bt.n1, M, Dm, m.qkv[i], epi(m.bqkv_scaled + static_cast<long long>(i) * 3 * Dm, bt.qkv), false, st));
            DAD_TRY(attention(bt.qkv, bt.att, bf, B, T, heads, st));
            DAD_TRY(m.linear(mode, bt.att, M, Dm, m.proj[i], epi(m.P(b + "attn.proj.bias"), bt.y1), false, st));
            DAD_TRY(ls_residual(bt.x0, bt.y1, bf, m.P(b + "ls1.gamma"), bt.x1, M, Dm, st));
            DAD_TRY(layernorm(bt.x1, m.P(b + "norm2.weight"), m.P(b + "norm2.bias"), bt.n2, bf, nullptr, M, Dm, 1, 1, 0, LN_EPS, st));
            if (Hd) {   // SwiGLUFFN (swiglu_ffn.py:30-34)
                DAD_TRY(m.linear(mode, bt.n2, M, Dm, m.w12[i], epi(m.P(b + "mlp.w12.bias"), bt.hpre), false, st));
                DAD_TRY(swiglu(bt.hpre, bt.h, bf, M, Hd, st));
                DAD_TRY(m.linear(mode, bt.h, M, Hd, m.w3[i], epi(m.P(b + "mlp.w3.bias"), bt.y2), false, st));
            } else {
                DAD_TRY(m.linear(mode, bt.n2, M, Dm, m.fc1[i], epi(m.P(b + "mlp.fc1.bias"), bt.hpre), false, st));
                DAD_TRY(gelu_fwd(bt.hpre, bt.h, bf, M * 4 * Dm, st));
                DAD_TRY(m.linear(mode, bt.h, M, 4 * Dm, m.fc2[i], epi(m.P(b + "mlp.fc2.bias"), bt.y2), false, st));
            }
            DAD_TRY(ls_residual(bt.x1, bt.y2, bf, m.P(b + "ls2.gamma"), xnext, M, Dm, st));
            if (tj < 4 && i == m.desc.taps[tj]) {
                DAD_TRY(layernorm(xnext, m.P(p + "norm.weight"), m.P(p + "norm.bias"), readout ? t.tapraw[tj] : t.tap[tj], bf,
                                  (tj == 3) ? feat_out : nullptr, Mp, Dm, np, T, 1, LN_EPS, st));
                if (readout) {   // tap = GELU(Linear(cat(patch, cls.expand_as(patch))))  (dpt.py:153-156)
                    const std::string r = "depth_head.readout_projects." + std::to_string(tj) + ".0.";
                    DAD_TRY(layernorm(xnext, m.P(p + "norm.weight"), m.P(p + "norm.bias"), t.cls[tj], bf, nullptr, B, Dm, 1, T, 0,
                                      LN_EPS, st));
                    DAD_TRY(concat_cls(t.tapraw[tj], t.cls[tj], t.cat[tj], bf, B, np, Dm, st));
                    DAD_TRY(m.linear(mode, t.cat[tj], Mp, 2 * Dm, m.readout_proj[tj], epi(m.P(r + "bias"), t.rpre[tj]), false, st));
                    DAD_TRY(gelu_fwd(t.rpre[tj], t.tap[tj], bf, Mp * Dm, st));
                }
                ++tj;
            }
        }
        DAD_REQUIRE(tj == 4, "taps must be increasing block indices < depth");

        const std::string h = "depth_head.", s = h + "scratch.";
        for (int j = 0; j < 4; ++j) {
            DAD_TRY(m.linear(mode, t.tap[j], Mp, Dm, m.projects[j], epi(m.P(h + "projects." + std::to_string(j) + ".bias"), t.pj[j]),
                             false, st));
            if (j == 0 || j == 1) {
                Epilogue es_ = epi(m.P(h + "resize_layers." + std::to_string(j) + ".bias"), t.rj[j]);
                es_.ldc = oc[j]; es_.scat_k = j == 0 ? 4 : 2; es_.scat_CoP = j == 0 ? m.CoP0 : m.CoP1; es_.scat_Co = oc[j];
                es_.scat_H = ph; es_.scat_W = pw;
                if (mode == 0 && oc[j] % 64 == 0)
                    DAD_TRY(m.conv(mode, t.pj[j], B, ph, pw, oc[j], j == 0 ? m.resize0 : m.resize1, 1, es_, false, st));
                else
                    DAD_TRY(m.linear(mode, t.pj[j], Mp, oc[j], j == 0 ? m.resize0 : m.resize1, es_, false, st));
            } else if (j == 3) {
                Epilogue e3 = epi(m.P(h + "resize_layers.3.bias"), t.rj[3]);
                if (mode == 0) {
                    DAD_TRY(m.conv(mode, t.pj[3], B, ph, pw, oc[3], m.resize3, 9, e3, false, st, 2));
                } else {
                    DAD_TRY(im2col_s2(t.pj[3], col, 0, B, ph, pw, oc[3], Cp3, st));
                    DAD_TRY(m.linear(mode, col, rows3, 9 * Cp3, m.resize3, e3, false, st));
                }
            }
            Epilogue er = epi(nullptr, t.lrn[j]); er.out_relu = t.lrn_relu[j];
            DAD_TRY(m.conv(mode, t.rj[j], B, t.hs[j], t.wsz[j], oc[j], m.layer_rn[j], 9, er, false, st));
        }
        for (int r = 3; r >= 0; --r) {
            FusionTape& f = t.fu[r];
            const std::string q = s + "refinenet" + std::to_string(r + 1) + ".";
            if (f.has_path) {
                Epilogue e1 = epi(m.P(q + "resConfUnit1.conv1.bias"), f.t1a); e1.act = ACT_RELU;
                DAD_TRY(m.conv(mode, f.lat_relu, B, f.H, f.W, F, m.rcu[r][0][0], 9, e1, false, st));
                Epilogue e2 = epi(m.P(q + "resConfUnit1.conv2.bias"), f.s); e2.res1 = f.lat; e2.res1_bf16 = bf; e2.res2 = f.path;
                e2.res2_bf16 = bf; e2.out_relu = f.sr;
                DAD_TRY(m.conv(mode, f.t1a, B, f.H, f.W, F, m.rcu[r][0][1], 9, e2, false, st));
            }
            Epilogue e1 = epi(m.P(q + "resConfUnit2.conv1.bias"), f.t1b); e1.act = ACT_RELU;
            DAD_TRY(m.conv(mode, f.sr, B, f.H, f.W, F, m.rcu[r][1][0], 9, e1, false, st));
            Epilogue e2 = epi(m.P(q + "resConfUnit2.conv2.bias"), f.u); e2.res1 = f.s; e2.res1_bf16 = bf;
            DAD_TRY(m.conv(mode, f.t1b, B, f.H, f.W, F, m.rcu[r][1][1], 9, e2, false, st));
            DAD_TRY(bilinear_nhwc(f.u, f.tmp, bf, B, f.H, f.W, f.Ho, f.Wo, F, st));   // reference order in both modes
            DAD_TRY(m.conv(mode, f.tmp, B, f.Ho, f.Wo, F, m.out_conv[r], 1, epi(m.P(q + "out_conv.bias"), f.res), false, st));
        }
        const int H1 = 2 * t.hs[0], W1 = 2 * t.wsz[0], F2 = F / 2;
        DAD_TRY(m.conv(mode, t.fu[0].res, B, H1, W1, F, m.output_conv1, 9, epi(m.P(s + "output_conv1.bias"), t.o1), false, st));
        DAD_TRY(bilinear_nhwc(t.o1, t.up, bf, B, H1, W1, H, W, F2, st));
        Epilogue eh = epi(m.P(s + "output_conv2.0.bias"), t.t32); eh.act = ACT_RELU;
        DAD_TRY(m.conv(mode, t.up, B, H, W, F2, m.output_conv2_0, 9, eh, false, st));
        const long long P = static_cast<long long>(B) * H * W;
        DAD_TRY(head1x1_any(t.t32, bf, m.P(s + "output_conv2.2.weight"), m.P(s + "output_conv2.2.bias"), t.depth, P, st));
        DAD_CHECK_CUDA(cudaMemcpyAsync(depth_out, t.depth, P * 4, cudaMemcpyDeviceToDevice, st));
        return DAD_OK;
    }

    // ------------------------------------------------------------------------------------ backward helpers
    static int rup(long long v, int q) { return static_cast<int>((v + q - 1) / q * q); }
    int ksplit_for(int Mrows, int Ncols, long long K) const {
        const long long tiles = static_cast<long long>(cdiv(Mrows, 128)) * cdiv(Ncols, 128);
        const long long want = cdivl(2LL * num_sms(), tiles);
        const long long maxs = std::max<long long>(1, cdivl(K, 64) / 4);   // at least 4 k-blocks per slice
        return static_cast<int>(std::max<long long>(1, std::min(want, maxs)));
    }
    // Weight-gradient operand path (A/B switch DAD_WGRAD_PATH): 2 (default) MN-major operands read in the activations' own
    // layouts (no copies), 1 channel-major copies over zero-padded pixel space (3x3 convolutions; linear as 0), 0 transposed
    // copies + the materialised 9x im2col^T operand.
    static int wgrad_path() {
        static const int v = [] { const char* e = getenv("DAD_WGRAD_PATH"); return e ? atoi(e) : 2; }();
        return v;
    }
    // out[Mo, No] (fp32) += At[Mo, K] Bt[No, K]^T on the tensor cores, split over K, partial tiles reduce-added through TMA
    int tc_accumulate(const void* At, const void* Bt, int Mo, int No, long long K, int Kp, float* out) {
        DAD_REQUIRE(No <= 16384, "backward: weight-gradient width %d exceeds the epilogue vectors", No);
        GemmProblem p;
        p.A = At; p.M = Mo; p.K = static_cast<int>(K); p.lda = Kp; p.Wt = Bt; p.N = No; p.Kp = Kp;
        p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = out; p.epi.out = out; p.epi.ldc = No;
        p.ksplit = std::max(2, ksplit_for(Mo, No, K));   // >= 2 keeps the problem on the split-K (1-CTA, reduce-add) kernel
        return gemm_tc(p, st);
    }

    // dW[Nout, Kin] += dY^T X        (dY [rows, Nout], X [rows, Kin], activation type)
    int wgrad_linear(const void* dY, long long ldy, const void* X, long long ldx, long long rows, int Nout, int Kin, float* dW,
                     Bump& ar) {
        if (!dW && !dry) return DAD_OK;   // the dry run sizes the scratch for the all-parameters case
        if (mode == 1) {
            if (dry) return DAD_OK;
            SGemm g; g.A = reinterpret_cast<const float*>(dY); g.sam = 1; g.sak = ldy; g.B = reinterpret_cast<const float*>(X);
            g.sbk = ldx; g.sbn = 1; g.C = dW; g.scm = Kin; g.scn = 1;
            g.M = Nout; g.N = Kin; g.K = static_cast<int>(rows); g.accumulate = 1;
            return sgemm(g, st);
        }
        if (wgrad_path() == 2 && ldy % 8 == 0 && ldx % 8 == 0 && Kin % 8 == 0) {
            // dY [rows, Nout] and X [rows, Kin] ARE the MN-major operands of dW = dY^T X: no transposed copies
            if (dry) return DAD_OK;
            DAD_REQUIRE(Kin <= 16384, "backward: weight-gradient width %d exceeds the epilogue vectors", Kin);
            GemmProblem p;
            p.mn = 1; p.A = dY; p.M = Nout; p.lda = ldy; p.Wt = X; p.N = Kin; p.ldw = ldx; p.K = static_cast<int>(rows);
            p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = dW; p.epi.out = dW; p.epi.ldc = Kin;
            p.ksplit = std::max(2, ksplit_for(Nout, Kin, rows));
            return gemm_tc(p, st);
        }
        const size_t mk = ar.used;
        const int Rp = rup(rows, 64);
        void* dYt = ar.bytes(static_cast<size_t>(Nout) * Rp * 2);
        void* Xt = ar.bytes(static_cast<size_t>(Kin) * Rp * 2);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_TRY(transpose_pad(dY, ldy, static_cast<int>(rows), Nout, dYt, Rp, st));
            DAD_TRY(transpose_pad(X, ldx, static_cast<int>(rows), Kin, Xt, Rp, st));
            DAD_TRY(tc_accumulate(dYt, Xt, Nout, Kin, rows, Rp, dW));
        }
        ar.used = mk;
        return DAD_OK;
    }
    // dX[rows, Kin] = dY[rows, Nout] W[Nout, Kin]   (W: fp32 master weight)
    int dgrad_linear(const void* dY, long long ldy, long long rows, int Nout, const float* Wm, int Kin, void* dX, Bump& ar) {
        if (mode == 1) {
            if (dry) return DAD_OK;
            SGemm g; g.A = reinterpret_cast<const float*>(dY); g.sam = ldy; g.sak = 1; g.B = Wm; g.sbk = Kin; g.sbn = 1;
            g.C = reinterpret_cast<float*>(dX); g.scm = Kin; g.scn = 1;
            g.M = static_cast<int>(rows); g.N = Kin; g.K = Nout;
            return sgemm(g, st);
        }
        const size_t mk = ar.used;
        const int Np = rup(Nout, 8);
        void* WT = ar.bytes(static_cast<size_t>(Kin) * Np * 2);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_TRY(pack_linear_T(Wm, WT, Nout, Kin, Np, st));
            GemmProblem p;
            p.A = dY; p.M = static_cast<int>(rows); p.K = Nout; p.lda = ldy; p.Wt = WT; p.N = Kin; p.Kp = Np;
            p.epi.bias = zeros; p.epi.out = dX; p.epi.out_bf16 = 1; p.epi.ldc = Kin;
            debug_label("gemm_tc dgrad");
            DAD_TRY(gemm_tc(p, st));
        }
        ar.used = mk;
        return DAD_OK;
    }
    int bias_grad(const void* dY, long long ld, long long rows, int N, float* db) {
        if (!db || dry) return DAD_OK;
        return colsum(dY, bf, ld, nullptr, 0, 0, rows, N, db, nullptr, nullptr, st);
    }
    // dW[Co, Ci, taps] += sum_pixels dOut[p, co] * window(X)[p, tap, ci];  X is [B, Hin, Win, Ci], dOut [B, Ho, Wo, Co]
    int conv_wgrad(const void* X, const void* dOut, int Hin, int Win, int Ci, int Co, int taps, int stride, int Ho, int Wo,
                   float* dW, Bump& ar) {
        if (!dW && !dry) return DAD_OK;   // the dry run sizes the scratch for the all-parameters case
        const long long P = static_cast<long long>(B) * Ho * Wo;
        if (mode == 1) {
            if (dry) return DAD_OK;
            SGemm g; g.A = reinterpret_cast<const float*>(dOut); g.sam = 1; g.sak = Co; g.B = reinterpret_cast<const float*>(X); g.C = dW;
            g.M = Co; g.N = taps * Ci; g.K = static_cast<int>(P); g.accumulate = 1;
            g.conv_taps = taps; g.convC = Ci; g.convH = Hin; g.convW = Win; g.convHo = Ho; g.convWo = Wo; g.conv_stride = stride;
            g.cmap = 1;
            return sgemm(g, st);
        }
        const size_t mk = ar.used;
        if (taps == 1 && stride == 1 && Ho == Hin && Wo == Win && wgrad_path() == 2 && Ci % 8 == 0 && Co % 8 == 0)
            return wgrad_linear(dOut, Co, X, Ci, P, Co, Ci, dW, ar);   // a 1x1 convolution is a linear layer over the pixels
        if (taps == 9 && stride == 1 && Ho == Hin && Wo == Win && wgrad_path() == 2 && Ci % 8 == 0 && Co % 8 == 0) {
            // 3x3 / stride 1, MN-major: dY and X (NHWC) are read as they are; the contraction runs over 8 x 8 pixel patches,
            // tap (dy, dx) is the same TMA box moved by (dy - 1, dx - 1) pixels with out-of-image pixels zero-filled.
            const int CiP = rup(Ci, 128);
            float* S = ar.f(static_cast<size_t>(Co) * 9 * CiP);
            if (!dry) {
                DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                DAD_REQUIRE(9 * CiP <= 16384, "backward: weight-gradient width %d exceeds the epilogue vectors", 9 * CiP);
                DAD_CHECK_CUDA(cudaMemsetAsync(S, 0, static_cast<size_t>(Co) * 9 * CiP * 4, st));
                GemmProblem p;
                p.mn = 2; p.A = dOut; p.M = Co; p.lda = Co; p.Wt = X; p.ldw = Ci; p.B = B; p.H = Hin; p.W = Win;
                p.shift_rows = Ci; p.shift_ld = CiP; p.N = 9 * CiP;
                p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = S; p.epi.out = S; p.epi.ldc = 9 * CiP;
                p.ksplit = std::max(2, ksplit_for(Co, 9 * CiP, static_cast<long long>(B) * cdiv(Hin, 8) * cdiv(Win, 8) * 64));
                DAD_TRY(gemm_tc(p, st));
                DAD_TRY(wgrad_unshift(S, dW, Co, Ci, 9, CiP, st));
            }
            ar.used = mk;
            return DAD_OK;
        }
        if (taps == 9 && stride == 1 && Ho == Hin && Wo == Win && wgrad_path() >= 1) {
            // 3x3 / stride 1: no 9x im2col operand.  Both operands are channel-major copies over ZERO-PADDED pixel space
            // q = (b, y + 1, x + 1) of an (H + 2) x Wp frame (Wp = W + 2 rounded up to 8): there a vertical tap is the constant
            // offset (dy - 1) * Wp, which TMA applies as a shifted K coordinate of ONE matrix (GemmProblem::shift_*); the
            // horizontal taps would be offsets of +-1 element, which TMA cannot fetch (the box start must be 16-byte aligned),
            // so X is written three times, pre-shifted by dx.  dY is zero on the frame, which removes every product the
            // convolution's zero padding does not contain.  The GEMM result is (tap, ci)-ordered; wgrad_unshift adds it
            // into the [Co, Ci, 3, 3] gradient.
            const int Wp = rup(Win + 2, 8);
            const long long Q = static_cast<long long>(B) * (Hin + 2) * Wp;
            const int Qp = rup(Q, 64);
            const int CiP = rup(Ci, 128);
            void* dOt = ar.bytes(static_cast<size_t>(Co) * Qp * 2);
            bf16* Xt = reinterpret_cast<bf16*>(ar.bytes(static_cast<size_t>(3) * CiP * Qp * 2));
            float* S = ar.f(static_cast<size_t>(Co) * 9 * CiP);
            if (!dry) {
                DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                DAD_REQUIRE(9 * CiP <= 16384, "backward: weight-gradient width %d exceeds the epilogue vectors", 9 * CiP);
                DAD_TRY(im2colT(dOut, B, Hin, Win, Co, 1, 1, Hin + 2, Wp, dOt, Qp, st, 1, 1));
                for (int dx = 0; dx < 3; ++dx) {   // block dx, frame pixel (oy, ox) = X(oy - 1, ox - 1 + dx - 1)
                    bf16* blk = Xt + static_cast<size_t>(dx) * CiP * Qp;
                    DAD_TRY(im2colT(X, B, Hin, Win, Ci, 1, 1, Hin + 2, Wp, blk, Qp, st, 1, 2 - dx));
                    if (CiP > Ci) DAD_CHECK_CUDA(cudaMemsetAsync(blk + static_cast<size_t>(Ci) * Qp, 0, static_cast<size_t>(CiP - Ci) * Qp * 2, st));
                }
                DAD_CHECK_CUDA(cudaMemsetAsync(S, 0, static_cast<size_t>(Co) * 9 * CiP * 4, st));
                GemmProblem p;
                p.A = dOt; p.M = Co; p.K = static_cast<int>(Q); p.lda = Qp; p.Wt = Xt; p.N = 9 * CiP; p.Kp = Qp;
                p.shift_taps = 9; p.shift_rows = 3 * CiP; p.shift_ld = CiP;
                for (int t = 0; t < 9; ++t) {
                    p.shift_off[t] = (t / 3 - 1) * Wp;
                    p.shift_row[t] = (t % 3) * CiP;
                }
                p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = S; p.epi.out = S; p.epi.ldc = 9 * CiP;
                p.ksplit = std::max(2, ksplit_for(Co, 9 * CiP, Q));
                DAD_TRY(gemm_tc(p, st));
                DAD_TRY(wgrad_unshift(S, dW, Co, Ci, 9, CiP, st));
            }
            ar.used = mk;
            return DAD_OK;
        }
        // rows of the im2col^T operand are ordered (ci, tap), so the GEMM output [Co][Ci*taps] IS the weight layout
        const int Pp = rup(P, 64);
        void* dOt = ar.bytes(static_cast<size_t>(Co) * Pp * 2);
        void* Xc = ar.bytes(static_cast<size_t>(Ci) * taps * Pp * 2);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_TRY(transpose_pad(dOut, Co, static_cast<int>(P), Co, dOt, Pp, st));
            DAD_TRY(im2colT(X, B, Hin, Win, Ci, taps, stride, Ho, Wo, Xc, Pp, st));
            DAD_TRY(tc_accumulate(dOt, Xc, Co, Ci * taps, P, Pp, dW));
        }
        ar.used = mk;
        return DAD_OK;
    }
    // stride-1 conv data gradient through the forward conv engine with flipped / transposed weights:
    // dIn[B,Hc,Wc,Ci] = conv(dOut[B,Hc,Wc,Co], Wd)
    int conv_dgrad(const void* dOut, int Hc, int Wc, int Co, int Ci, int taps, const float* Wmaster, void* dIn, Bump& ar) {
        const int CoP = cdiv(Co, 64) * 64;
        const size_t mk = ar.used;
        void* wd = ar.bytes(static_cast<size_t>(Ci) * taps * CoP * es);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_TRY(pack_conv_dgrad(Wmaster, wd, bf, Co, Ci, taps, CoP, st));
            Mat mt; mt.w[mode] = wd; mt.N = Ci; mt.Kp = taps * CoP;
            DAD_TRY(m.conv(mode, dOut, B, Hc, Wc, Co, mt, taps, epi(nullptr, dIn), false, st));
        }
        ar.used = mk;  // stream order keeps wd alive until the conv has read it; the next user writes after it
        return DAD_OK;
    }
    // gin (activation type, [B,Hi,Wi,C]) = adjoint of the bilinear resampling applied to gout [B,Ho,Wo,C]
    int bilinear_adjoint(const void* gout, void* gin, int Hi, int Wi, int Ho, int Wo, int C, Bump& ar) {
        const long long n = static_cast<long long>(B) * Hi * Wi * C;
        static const bool scatter = getenv("DAD_BILINEAR_SCATTER") != nullptr;   // A/B switch: round 1's atomicAdd form
        if (!scatter) return dry ? DAD_OK : bilinear_bwd_gather(gout, bf, gin, B, Hi, Wi, Ho, Wo, C, st);
        const size_t mk = ar.used;
        float* acc = bf ? ar.f(n) : reinterpret_cast<float*>(gin);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_CHECK_CUDA(cudaMemsetAsync(acc, 0, n * 4, st));
            DAD_TRY(bilinear_bwd(gout, bf, acc, B, Hi, Wi, Ho, Wo, C, st));
            if (bf) DAD_TRY(convert(acc, 0, gin, 1, n, st));
        }
        ar.used = mk;
        return DAD_OK;
    }
    // fp32 view of an activation tensor (mode 0: converted copy in scratch; mode 1: the tensor itself)
    const float* as_f32(const void* src, long long n, Bump& ar) {
        if (!bf) return reinterpret_cast<const float*>(src);
        float* d = ar.f(n);
        if (!dry && !ar.overflow) convert(src, 1, d, 0, n, st);
        return d;
    }

    // one batched tcgen05 launch: per (image, head) out = A W^T over K-major bf16 operands (see GemmProblem::batch_h)
    int batched_tc(const void* A, long long lda, long long a_sh, long long a_sb, const void* Wt, long long ldw, long long w_sh,
                   long long w_sb, int w_rows, int Mrows, int N, int K, void* out, int out_bf16, long long ldc, long long c_row_b,
                   long long c_row_h, int c_col_h, const float* scale) {
        GemmProblem p;
        p.A = A; p.M = Mrows; p.K = K; p.lda = lda; p.a_sh = a_sh; p.a_sb = a_sb;
        p.Wt = Wt; p.N = N; p.Kp = rup(K, 8); p.ldw = ldw; p.w_sh = w_sh; p.w_sb = w_sb; p.w_rows = w_rows;
        p.batch_h = heads; p.batch_b = B;
        p.c_row_b = c_row_b; p.c_row_h = c_row_h; p.c_col_h = c_col_h;
        p.epi.out = out; p.epi.out_bf16 = out_bf16; p.epi.ldc = ldc; p.epi.gamma = scale;
        return gemm_tc(p, st);
    }

    // bf16 engine: the five contractions of the attention backward as batched tcgen05 GEMMs (one problem per image and
    // head) over K-major operands.  S and dP are materialised in fp32 ([Z][T][Tp], Tp = T rounded up to 128) by the TMA
    // reduce-add epilogue onto zero-filled buffers; the row-wise softmax kernels turn them into bf16 P / dS, which the
    // remaining three GEMMs read directly or through 64 x 64 tile transposes.
    int attention_bwd_tc(const BlockTape& bt, const void* datt, void* dqkv, Bump& ar) {
        const size_t mk = ar.used;
        const int Tp = rup(T, 128), Z = B * heads;
        const long long tt = static_cast<long long>(T) * Tp, sq = static_cast<long long>(Tp) * Tp, ld = 3LL * Dm;
        float* S = ar.f(static_cast<size_t>(Z) * tt);
        float* dPf = ar.f(static_cast<size_t>(Z) * tt);
        void* Pn = ar.bytes(static_cast<size_t>(Z) * tt * 2);    // P  [z][i][j]
        void* dS = ar.bytes(static_cast<size_t>(Z) * tt * 2);    // dS [z][i][j]
        // MN-major operands (default): P / dS, q / k / v and dO are read where they lie - a contraction over the rows of a
        // row-major matrix IS an MN-major operand - so the transposed copies below exist only on the A/B path
        const bool mnp = wgrad_path() == 2;
        void* Pt = mnp ? nullptr : ar.bytes(static_cast<size_t>(Z) * sq * 2);    // P  [z][j][i]
        void* dSt = mnp ? nullptr : ar.bytes(static_cast<size_t>(Z) * sq * 2);   // dS [z][j][i]
        void* Qt = mnp ? nullptr : ar.bytes(static_cast<size_t>(Z) * 64 * Tp * 2);
        void* Kt = mnp ? nullptr : ar.bytes(static_cast<size_t>(Z) * 64 * Tp * 2);
        void* dOt = mnp ? nullptr : ar.bytes(static_cast<size_t>(Z) * 64 * Tp * 2);
        if (!dry && mnp) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            const bf16* q = reinterpret_cast<const bf16*>(bt.qkv);
            bf16* dq = reinterpret_cast<bf16*>(dqkv);
            const long long zrows = static_cast<long long>(Z) * T;
            auto scores = [&](const void* A, long long lda, long long a_sb, const void* Wt, float* out) -> int {
                GemmProblem p;   // plain clipped stores (c_store): no zero fill of the 2 x Z x T x Tp fp32 destinations, no read
                p.c_store = 1;
                p.A = A; p.M = T; p.K = 64; p.lda = lda; p.a_sh = 64; p.a_sb = a_sb;
                p.Wt = Wt; p.N = Tp; p.Kp = 64; p.ldw = ld; p.w_sh = 64; p.w_sb = T * ld; p.w_rows = T;
                p.batch_h = heads; p.batch_b = B; p.c_row_b = static_cast<long long>(heads) * T; p.c_row_h = T;
                p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = out; p.epi.out = out; p.epi.ldc = Tp;
                return gemm_tc(p, st);
            };
            // out[z][row, 0:64] (a 64-column slice of dqkv) = scale * op(A)[z] W[z]: A = P / dS [z][T][Tp] read MN-major (mn 1:
            // the contraction runs over its rows) or K-major (mn 3), W = a [T][64] head slice of q / k / dO read MN-major
            auto apply = [&](int mn, const void* A, const bf16* Wm, long long ldw, bf16* out, const float* scale) -> int {
                GemmProblem p;
                p.mn = mn; p.A = A; p.M = T; p.K = T; p.lda = Tp; p.a_sh = tt; p.a_sb = heads * tt;
                p.Wt = Wm; p.N = 64; p.ldw = ldw; p.w_sh = 64; p.w_sb = T * ldw; p.w_rows = 64;
                p.batch_h = heads; p.batch_b = B; p.c_row_b = T; p.c_row_h = 0; p.c_col_h = 64;
                p.epi.out = out; p.epi.out_bf16 = 1; p.epi.ldc = ld; p.epi.gamma = scale;
                return gemm_tc(p, st);
            };
            DAD_TRY(scores(q, ld, T * ld, q + Dm, S));                                           // S  = Q' K^T
            DAD_TRY(scores(datt, Dm, static_cast<long long>(T) * Dm, q + 2 * Dm, dPf));          // dP = dO V^T
            DAD_TRY(softmax_rows_bf16(S, Pn, zrows, T, Tp, st));
            DAD_TRY(apply(1, Pn, reinterpret_cast<const bf16*>(datt), Dm, dq + 2 * Dm, nullptr));   // dV = P^T dO
            DAD_TRY(softmax_bwd_bf16(Pn, dPf, dS, zrows, T, Tp, st));
            DAD_TRY(apply(3, dS, q + Dm, ld, dq, eighths));                                      // dq = 0.125 dS K
            DAD_TRY(apply(1, dS, q, ld, dq + Dm, nullptr));                                      // dK = dS^T Q'
        }
        if (!dry && !mnp) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            const bf16* q = reinterpret_cast<const bf16*>(bt.qkv);
            bf16* dq = reinterpret_cast<bf16*>(dqkv);
            const long long zrows = static_cast<long long>(Z) * T;
            auto scores = [&](const void* A, long long lda, long long a_sb, const void* Wt, float* out) -> int {
                DAD_CHECK_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(Z) * tt * 4, st));
                GemmProblem p;
                p.A = A; p.M = T; p.K = 64; p.lda = lda; p.a_sh = 64; p.a_sb = a_sb;
                p.Wt = Wt; p.N = Tp; p.Kp = 64; p.ldw = ld; p.w_sh = 64; p.w_sb = T * ld; p.w_rows = T;
                p.batch_h = heads; p.batch_b = B; p.c_row_b = static_cast<long long>(heads) * T; p.c_row_h = T;
                p.epi.bias = zeros; p.epi.gamma = ones; p.epi.res1 = out; p.epi.out = out; p.epi.ldc = Tp;
                return gemm_tc(p, st);
            };
            DAD_TRY(scores(q, ld, T * ld, q + Dm, S));                                           // S  = Q' K^T
            DAD_TRY(scores(datt, Dm, static_cast<long long>(T) * Dm, q + 2 * Dm, dPf));          // dP = dO V^T
            DAD_TRY(softmax_rows_bf16(S, Pn, zrows, T, Tp, st));
            DAD_TRY(transpose_pad_batched(Pn, Tp, T, Tp, Pt, Tp, Z, tt, sq, st));
            DAD_TRY(head_transpose(datt, Dm, dOt, B, T, heads, Tp, st));
            DAD_TRY(head_transpose(q, ld, Qt, B, T, heads, Tp, st));
            DAD_TRY(head_transpose(q + Dm, ld, Kt, B, T, heads, Tp, st));
            // dV[j, d] = sum_i P[i, j] dO[i, d]
            DAD_TRY(batched_tc(Pt, Tp, sq, heads * sq, dOt, Tp, 64LL * Tp, heads * 64LL * Tp, 64, T, 64, T, dq + 2 * Dm, 1, ld, T, 0, 64,
                               nullptr));
            DAD_TRY(softmax_bwd_bf16(Pn, dPf, dS, zrows, T, Tp, st));
            DAD_TRY(transpose_pad_batched(dS, Tp, T, Tp, dSt, Tp, Z, tt, sq, st));
            // dq = 0.125 * dS K  (the packed q rows carry 64^-0.5), dK = dS^T Q'
            DAD_TRY(batched_tc(dS, Tp, tt, heads * tt, Kt, Tp, 64LL * Tp, heads * 64LL * Tp, 64, T, 64, T, dq, 1, ld, T, 0, 64, eighths));
            DAD_TRY(batched_tc(dSt, Tp, sq, heads * sq, Qt, Tp, 64LL * Tp, heads * 64LL * Tp, 64, T, 64, T, dq + Dm, 1, ld, T, 0, 64,
                               nullptr));
        }
        ar.used = mk;
        return DAD_OK;
    }

    int attention_bwd(const BlockTape& bt, const void* datt_any, void* dqkv_any, Bump& ar) {
        if (bf && !m.attn_bwd_fp32) return attention_bwd_tc(bt, datt_any, dqkv_any, ar);
        const size_t mk = ar.used;
        const long long tt = static_cast<long long>(T) * T;
        const float* qkv = as_f32(bt.qkv, M * 3 * Dm, ar);
        const float* datt = as_f32(datt_any, M * Dm, ar);
        float* dqkv = bf ? ar.f(M * 3 * Dm) : reinterpret_cast<float*>(dqkv_any);
        float* Pm = ar.f(static_cast<size_t>(B) * heads * tt);
        float* dP = ar.f(static_cast<size_t>(B) * heads * tt);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            const long long ld = 3LL * Dm;
            SGemm s;  // S = Q' K^T
            s.A = qkv; s.sam = ld; s.sak = 1; s.a1 = T * ld; s.a2 = 64;
            s.B = qkv + Dm; s.sbk = 1; s.sbn = ld; s.b1 = T * ld; s.b2 = 64;
            s.C = Pm; s.scm = T; s.scn = 1; s.c1 = heads * tt; s.c2 = tt;
            s.M = T; s.N = T; s.K = 64; s.nb1 = B; s.nb2 = heads;
            DAD_TRY(sgemm(s, st));
            DAD_TRY(softmax_rows(Pm, static_cast<long long>(B) * heads * T, T, T, st));
            SGemm p;  // dP = dO V^T
            p.A = datt; p.sam = Dm; p.sak = 1; p.a1 = static_cast<long long>(T) * Dm; p.a2 = 64;
            p.B = qkv + 2 * Dm; p.sbk = 1; p.sbn = ld; p.b1 = T * ld; p.b2 = 64;
            p.C = dP; p.scm = T; p.scn = 1; p.c1 = heads * tt; p.c2 = tt;
            p.M = T; p.N = T; p.K = 64; p.nb1 = B; p.nb2 = heads;
            DAD_TRY(sgemm(p, st));
            SGemm v;  // dV = P^T dO
            v.A = Pm; v.sam = 1; v.sak = T; v.a1 = heads * tt; v.a2 = tt;
            v.B = datt; v.sbk = Dm; v.sbn = 1; v.b1 = static_cast<long long>(T) * Dm; v.b2 = 64;
            v.C = dqkv + 2 * Dm; v.scm = ld; v.scn = 1; v.c1 = T * ld; v.c2 = 64;
            v.M = T; v.N = 64; v.K = T; v.nb1 = B; v.nb2 = heads;
            DAD_TRY(sgemm(v, st));
            DAD_TRY(softmax_bwd_rows(Pm, dP, static_cast<long long>(B) * heads * T, T, T, st));   // dP <- dS
            SGemm q;  // dq = 0.125 * dS K   (the packed q rows carry 64^-0.5: q' = q / 8)
            q.A = dP; q.sam = T; q.sak = 1; q.a1 = heads * tt; q.a2 = tt;
            q.B = qkv + Dm; q.sbk = ld; q.sbn = 1; q.b1 = T * ld; q.b2 = 64;
            q.C = dqkv; q.scm = ld; q.scn = 1; q.c1 = T * ld; q.c2 = 64;
            q.M = T; q.N = 64; q.K = T; q.nb1 = B; q.nb2 = heads; q.alpha = 0.125f;
            DAD_TRY(sgemm(q, st));
            SGemm k;  // dK = dS^T Q'
            k.A = dP; k.sam = 1; k.sak = T; k.a1 = heads * tt; k.a2 = tt;
            k.B = qkv; k.sbk = ld; k.sbn = 1; k.b1 = T * ld; k.b2 = 64;
            k.C = dqkv + Dm; k.scm = ld; k.scn = 1; k.c1 = T * ld; k.c2 = 64;
            k.M = T; k.N = 64; k.K = T; k.nb1 = B; k.nb2 = heads;
            DAD_TRY(sgemm(k, st));
            if (bf) DAD_TRY(convert(dqkv, 0, dqkv_any, 1, M * 3 * Dm, st));
        }
        ar.used = mk;
        return DAD_OK;
    }

    // one FeatureFusionBlock backward (util/blocks.py:129-146); dres [B,Ho,Wo,F] -> dlat [B,H,W,F], dpath (has_path)
    int fusion_bwd(int r, const FusionTape& f, const void* dres, void** dlat_out, void** dpath_out, Bump& ar) {
        const std::string q = "depth_head.scratch.refinenet" + std::to_string(r + 1) + ".";
        const long long n = static_cast<long long>(B) * f.H * f.W * F, no = static_cast<long long>(B) * f.Ho * f.Wo * F;
        const long long px = n / F, pxo = no / F;
        DAD_TRY(conv_wgrad(f.tmp, dres, f.Ho, f.Wo, F, F, 1, 1, f.Ho, f.Wo, G(q + "out_conv.weight"), ar));
        DAD_TRY(bias_grad(dres, F, pxo, F, G(q + "out_conv.bias")));
        void* dtmp = a(ar, no);
        DAD_TRY(conv_dgrad(dres, f.Ho, f.Wo, F, F, 1, m.P(q + "out_conv.weight"), dtmp, ar));
        void* du = a(ar, n);
        DAD_TRY(bilinear_adjoint(dtmp, du, f.H, f.W, f.Ho, f.Wo, F, ar));
        // RCU2: u = conv2(relu(conv1(relu(s)) + b1)) + b2 + s
        DAD_TRY(conv_wgrad(f.t1b, du, f.H, f.W, F, F, 9, 1, f.H, f.W, G(q + "resConfUnit2.conv2.weight"), ar));
        DAD_TRY(bias_grad(du, F, px, F, G(q + "resConfUnit2.conv2.bias")));
        void* dt1b = a(ar, n);
        DAD_TRY(conv_dgrad(du, f.H, f.W, F, F, 9, m.P(q + "resConfUnit2.conv2.weight"), dt1b, ar));
        RUN(relu_bwd(dt1b, f.t1b, nullptr, dt1b, bf, n, st));
        DAD_TRY(conv_wgrad(f.sr, dt1b, f.H, f.W, F, F, 9, 1, f.H, f.W, G(q + "resConfUnit2.conv1.weight"), ar));
        DAD_TRY(bias_grad(dt1b, F, px, F, G(q + "resConfUnit2.conv1.bias")));
        void* ds = a(ar, n);
        DAD_TRY(conv_dgrad(dt1b, f.H, f.W, F, F, 9, m.P(q + "resConfUnit2.conv1.weight"), ds, ar));
        RUN(relu_bwd(ds, f.sr, du, ds, bf, n, st));   // ds = du + dsr * (s > 0)
        if (!f.has_path) {
            *dlat_out = ds;
            *dpath_out = nullptr;
            return DAD_OK;
        }
        // RCU1 on the lateral: s = conv2(relu(conv1(relu(lat)) + b1)) + b2 + lat + path
        DAD_TRY(conv_wgrad(f.t1a, ds, f.H, f.W, F, F, 9, 1, f.H, f.W, G(q + "resConfUnit1.conv2.weight"), ar));
        DAD_TRY(bias_grad(ds, F, px, F, G(q + "resConfUnit1.conv2.bias")));
        void* dt1a = a(ar, n);
        DAD_TRY(conv_dgrad(ds, f.H, f.W, F, F, 9, m.P(q + "resConfUnit1.conv2.weight"), dt1a, ar));
        RUN(relu_bwd(dt1a, f.t1a, nullptr, dt1a, bf, n, st));
        DAD_TRY(conv_wgrad(f.lat_relu, dt1a, f.H, f.W, F, F, 9, 1, f.H, f.W, G(q + "resConfUnit1.conv1.weight"), ar));
        DAD_TRY(bias_grad(dt1a, F, px, F, G(q + "resConfUnit1.conv1.bias")));
        void* dlat = a(ar, n);
        DAD_TRY(conv_dgrad(dt1a, f.H, f.W, F, F, 9, m.P(q + "resConfUnit1.conv1.weight"), dlat, ar));
        RUN(relu_bwd(dlat, f.lat_relu, ds, dlat, bf, n, st));   // dlat = ds + dlr * (lat > 0)
        *dlat_out = dlat;
        *dpath_out = ds;
        return DAD_OK;
    }

    // ------------------------------------------------------------------------------------ backward
    int backward(const float* gdepth, const float* gfeat, Bump& ar, Tape& t) {
        plan(ar, t);
        if (!dry) DAD_REQUIRE(gdepth, "backward: grad_depth must not be null");
        const std::string p = "pretrained.", h = "depth_head.", s = h + "scratch.";
        const int F2 = F / 2, H1 = 2 * t.hs[0], W1 = 2 * t.wsz[0];
        const long long P = static_cast<long long>(B) * H * W, P1 = static_cast<long long>(B) * H1 * W1;
        if (bf) {
            constexpr int VN = 16384;
            float* v = ar.f(3 * VN);
            if (!dry) {
                DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                DAD_TRY(fill_f32(v, 0.f, VN, st));
                DAD_TRY(fill_f32(v + VN, 1.f, VN, st));
                DAD_TRY(fill_f32(v + 2 * VN, 0.125f, VN, st));
            }
            zeros = v; ones = v ? v + VN : nullptr; eighths = v ? v + 2 * VN : nullptr;
        }

        // ---- output head
        void* dt32 = a(ar, P * 32);
        if (!dry) DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
        RUN(head_bwd(gdepth, t.depth, t.t32, bf, m.P(s + "output_conv2.2.weight"), dt32, G(s + "output_conv2.2.weight"),
                     G(s + "output_conv2.2.bias"), P, st));
        DAD_TRY(conv_wgrad(t.up, dt32, H, W, F2, 32, 9, 1, H, W, G(s + "output_conv2.0.weight"), ar));
        DAD_TRY(bias_grad(dt32, 32, P, 32, G(s + "output_conv2.0.bias")));
        void* dup = a(ar, P * F2);
        DAD_TRY(conv_dgrad(dt32, H, W, 32, F2, 9, m.P(s + "output_conv2.0.weight"), dup, ar));
        void* do1 = a(ar, P1 * F2);
        DAD_TRY(bilinear_adjoint(dup, do1, H1, W1, H, W, F2, ar));
        DAD_TRY(conv_wgrad(t.fu[0].res, do1, H1, W1, F, F2, 9, 1, H1, W1, G(s + "output_conv1.weight"), ar));
        DAD_TRY(bias_grad(do1, F2, P1, F2, G(s + "output_conv1.bias")));
        void* dres = a(ar, P1 * F);
        DAD_TRY(conv_dgrad(do1, H1, W1, F2, F, 9, m.P(s + "output_conv1.weight"), dres, ar));

        // ---- fusion blocks, finest first
        void* dlat[4];
        for (int r = 0; r < 4; ++r) {
            void* dpath = nullptr;
            DAD_TRY(fusion_bwd(r, t.fu[r], dres, &dlat[r], &dpath, ar));
            dres = dpath;
        }

        // ---- reassemble: layer_rn -> resize -> projects; dtap[j] = gradient of the LayerNorm'd tap
        void* dtap[4];
        float* dcls[4] = {nullptr, nullptr, nullptr, nullptr};
        for (int j = 0; j < 4; ++j) {
            const std::string js = std::to_string(j);
            const int hj = t.hs[j], wj = t.wsz[j];
            const long long pxj = static_cast<long long>(B) * hj * wj;
            DAD_TRY(conv_wgrad(t.rj[j], dlat[j], hj, wj, oc[j], F, 9, 1, hj, wj, G(s + "layer" + std::to_string(j + 1) + "_rn.weight"), ar));
            void* drj = a(ar, pxj * oc[j]);
            DAD_TRY(conv_dgrad(dlat[j], hj, wj, F, oc[j], 9, m.P(s + "layer" + std::to_string(j + 1) + "_rn.weight"), drj, ar));
            void* dpj = drj;
            if (j != 2 && bf) {
                // bf16 engine: the same contractions through the tensor-core helpers
                dpj = a(ar, Mp * oc[j]);
                DAD_TRY(bias_grad(drj, oc[j], pxj, oc[j], G(h + "resize_layers." + js + ".bias")));
                const size_t mk = ar.used;
                if (j == 3) {
                    const int Cp = cdiv(oc[3], 64) * 64, CoP8 = rup(oc[3], 8);
                    DAD_TRY(conv_wgrad(t.pj[3], drj, ph, pw, oc[3], oc[3], 9, 2, hj, wj, G(h + "resize_layers.3.weight"), ar));
                    void* WT = ar.bytes(static_cast<size_t>(9) * Cp * CoP8 * 2);
                    void* dcol = a(ar, pxj * 9 * Cp);
                    if (!dry) {
                        DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                        // dcol[pxj, 9*Cp] = drj[pxj, Co] W[Co, 9*Cp]  (the packed fp32 matrix of the forward), then col2im
                        DAD_TRY(pack_linear_T(reinterpret_cast<const float*>(m.resize3.w[1]), WT, oc[3], 9 * Cp, CoP8, st));
                        GemmProblem p;
                        p.A = drj; p.M = static_cast<int>(pxj); p.K = oc[3]; p.lda = oc[3]; p.Wt = WT; p.N = 9 * Cp; p.Kp = CoP8;
                        p.epi.bias = zeros; p.epi.out = dcol; p.epi.out_bf16 = 1; p.epi.ldc = 9 * Cp;
                        DAD_TRY(gemm_tc(p, st));
                        DAD_TRY(col2im_s2(dcol, dpj, 1, B, ph, pw, oc[3], Cp, st));
                    }
                } else {
                    const int k = j == 0 ? 4 : 2, kk = k * k, CoP = j == 0 ? m.CoP0 : m.CoP1;
                    const Mat& mt = j == 0 ? m.resize0 : m.resize1;
                    void* Gm = a(ar, Mp * kk * CoP);
                    float* tmpw = ar.f(static_cast<size_t>(kk) * CoP * oc[j]);
                    if (!dry) {
                        DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                        DAD_TRY(convT_gather(drj, Gm, 1, B, ph, pw, k, oc[j], CoP, st));
                    }
                    float* dW = G(h + "resize_layers." + js + ".weight");
                    if (dW || dry) {
                        if (!dry) DAD_CHECK_CUDA(cudaMemsetAsync(tmpw, 0, static_cast<size_t>(kk) * CoP * oc[j] * 4, st));
                        DAD_TRY(wgrad_linear(Gm, static_cast<long long>(kk) * CoP, t.pj[j], oc[j], Mp, kk * CoP, oc[j], tmpw, ar));
                        if (!dry) DAD_TRY(convT_wgrad_permute(tmpw, dW, oc[j], oc[j], CoP, kk, st));
                    }
                    // dpj[Mp, Ci] = Gm[Mp, kk*CoP] Wm[kk*CoP, Ci]  (the packed fp32 ConvTranspose matrix of the forward)
                    DAD_TRY(dgrad_linear(Gm, static_cast<long long>(kk) * CoP, Mp, kk * CoP, reinterpret_cast<const float*>(mt.w[1]), oc[j],
                                         dpj, ar));
                }
                ar.used = mk;
            } else if (j != 2) {
                // ConvTranspose (j = 0, 1) / stride-2 conv (j = 3): small GEMMs, fp32 engine in both modes
                dpj = a(ar, Mp * oc[j]);
                const size_t mk = ar.used;
                const float* drj_f = as_f32(drj, pxj * oc[j], ar);
                const float* pj_f = as_f32(t.pj[j], Mp * oc[j], ar);
                float* dpj_f = bf ? ar.f(Mp * oc[j]) : reinterpret_cast<float*>(dpj);
                RUN(colsum(drj_f, 0, oc[j], nullptr, 0, 0, pxj, oc[j], G(h + "resize_layers." + js + ".bias"), nullptr, nullptr, st));
                if (j == 3) {
                    const int Cp = cdiv(oc[3], 64) * 64;
                    float* dcol = ar.f(pxj * 9 * Cp);
                    if (!dry) {
                        DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                        if (float* dW = G(h + "resize_layers.3.weight")) {
                            SGemm g; g.A = drj_f; g.sam = 1; g.sak = oc[3]; g.B = pj_f; g.C = dW;
                            g.M = oc[3]; g.N = 9 * oc[3]; g.K = static_cast<int>(pxj); g.accumulate = 1;
                            g.conv_taps = 9; g.convC = oc[3]; g.convH = ph; g.convW = pw; g.convHo = hj; g.convWo = wj; g.conv_stride = 2;
                            g.cmap = 1;
                            DAD_TRY(sgemm(g, st));
                        }
                        SGemm d; d.A = drj_f; d.sam = oc[3]; d.sak = 1; d.B = reinterpret_cast<const float*>(m.resize3.w[1]); d.sbk = 9 * Cp;
                        d.sbn = 1; d.C = dcol; d.scm = 9 * Cp; d.scn = 1; d.M = static_cast<int>(pxj); d.N = 9 * Cp; d.K = oc[3];
                        DAD_TRY(sgemm(d, st));
                        DAD_TRY(col2im_s2(dcol, dpj_f, 0, B, ph, pw, oc[3], Cp, st));
                    }
                } else {
                    const int k = j == 0 ? 4 : 2, kk = k * k, CoP = j == 0 ? m.CoP0 : m.CoP1;
                    const Mat& mt = j == 0 ? m.resize0 : m.resize1;
                    float* Gm = ar.f(Mp * kk * CoP);
                    if (!dry) {
                        DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                        DAD_TRY(convT_gather(drj_f, Gm, 0, B, ph, pw, k, oc[j], CoP, st));
                        if (float* dW = G(h + "resize_layers." + js + ".weight")) {
                            SGemm g; g.A = Gm; g.sam = 1; g.sak = static_cast<long long>(kk) * CoP; g.B = pj_f; g.sbk = oc[j]; g.sbn = 1;
                            g.C = dW; g.M = kk * CoP; g.N = oc[j]; g.K = static_cast<int>(Mp); g.accumulate = 1;
                            g.cmap = 2; g.ct_CoP = CoP; g.ct_Co = oc[j]; g.ct_kk = kk;
                            DAD_TRY(sgemm(g, st));
                        }
                        // dpj[Mp, Ci] = Gm[Mp, kk*CoP] Wm[kk*CoP, Ci]  (the packed fp32 ConvTranspose matrix)
                        SGemm d; d.A = Gm; d.sam = static_cast<long long>(kk) * CoP; d.sak = 1; d.B = reinterpret_cast<const float*>(mt.w[1]);
                        d.sbk = mt.Kp; d.sbn = 1; d.C = dpj_f; d.scm = oc[j]; d.scn = 1; d.M = static_cast<int>(Mp); d.N = oc[j]; d.K = kk * CoP;
                        DAD_TRY(sgemm(d, st));
                    }
                }
                if (bf) RUN(convert(dpj_f, 0, dpj, 1, Mp * oc[j], st));
                ar.used = mk;
            }
            DAD_TRY(wgrad_linear(dpj, oc[j], t.tap[j], Dm, Mp, oc[j], Dm, G(h + "projects." + js + ".weight"), ar));
            DAD_TRY(bias_grad(dpj, oc[j], Mp, oc[j], G(h + "projects." + js + ".bias")));
            dtap[j] = a(ar, Mp * Dm);
            if (!dry) DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_TRY(dgrad_linear(dpj, oc[j], Mp, oc[j], m.P(h + "projects." + js + ".weight"), Dm, dtap[j], ar));
            if (readout) {
                // dtap[j] is the gradient of GELU(Linear(cat)); take it back to the normalised patch tokens (left half of
                // the concat, kept in dtap[j]) and to the class token (right half summed over the image's patches, fp32)
                const std::string r = h + "readout_projects." + js + ".0.";
                RUN(gelu_bwd(t.rpre[j], dtap[j], dtap[j], bf, Mp * Dm, st));
                DAD_TRY(wgrad_linear(dtap[j], Dm, t.cat[j], 2 * Dm, Mp, Dm, 2 * Dm, G(r + "weight"), ar));
                DAD_TRY(bias_grad(dtap[j], Dm, Mp, Dm, G(r + "bias")));
                dcls[j] = ar.f(static_cast<size_t>(B) * Dm);
                const size_t mk = ar.used;
                void* dcat = a(ar, Mp * 2 * Dm);
                if (!dry) DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                DAD_TRY(dgrad_linear(dtap[j], Dm, Mp, Dm, m.P(r + "weight"), 2 * Dm, dcat, ar));
                if (!dry) {
                    DAD_TRY(copy_cols(dcat, 2 * Dm, dtap[j], Dm, Mp, Dm, bf, st));
                    DAD_CHECK_CUDA(cudaMemsetAsync(dcls[j], 0, static_cast<size_t>(B) * Dm * 4, st));
                    for (int bi = 0; bi < B; ++bi) {
                        const uint8_t* right = reinterpret_cast<const uint8_t*>(dcat) +
                                               (static_cast<size_t>(bi) * np * 2 * Dm + Dm) * es;
                        DAD_TRY(colsum(right, bf, 2 * Dm, nullptr, 0, 0, np, Dm, dcls[j] + static_cast<size_t>(bi) * Dm, nullptr,
                                       nullptr, st));
                    }
                }
                ar.used = mk;
            }
        }
        if (gfeat) RUN(add_inplace(dtap[3], bf, gfeat, Mp * Dm, st));   // features[3][0]: the normalised patch tokens

        // ---- encoder, last block first.  Gx = gradient of the residual stream (fp32 in both modes)
        float* Gx = ar.f(M * Dm);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_CHECK_CUDA(cudaMemsetAsync(Gx, 0, M * Dm * 4, st));
        }
        int tj = 3;
        for (int i = L - 1; i >= 0; --i) {
            const std::string b = p + "blocks." + std::to_string(i) + ".";
            const BlockTape& bt = t.blk[i];
            const float* xnext = (i + 1 < L) ? t.blk[i + 1].x0 : t.xfinal;
            if (tj >= 0 && i == m.desc.taps[tj]) {
                RUN(layernorm_bwd(xnext, m.P(p + "norm.weight"), dtap[tj], bf, Gx, G(p + "norm.weight"), G(p + "norm.bias"), Mp, Dm, np,
                                  T, 1, LN_EPS, st));
                if (readout)   // the class-token rows of the same LayerNorm (input row b * T)
                    RUN(layernorm_bwd(xnext, m.P(p + "norm.weight"), dcls[tj], 0, Gx, G(p + "norm.weight"), G(p + "norm.bias"), B, Dm,
                                      1, T, 0, LN_EPS, st));
                --tj;
            }
            const size_t mk = ar.used;
            void* dy = a(ar, M * Dm);        // gradient of the branch output before LayerScale
            void* dh = a(ar, M * (Hd ? 2 * Hd : 4 * Dm));
            void* dg = Hd ? a(ar, M * Hd) : nullptr;   // SwiGLU: gradient of the gated product
            void* dn = a(ar, M * Dm);
            void* dqkv = a(ar, M * 3 * Dm);
            if (!dry) DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            // x_{i+1} = x1 + gamma2 * y2
            RUN(colsum(Gx, 0, Dm, bt.y2, bf, Dm, M, Dm, G(b + "ls2.gamma"), m.P(b + "ls2.gamma"), dy, st));
            if (Hd) {   // y2 = w3(silu(x1) * x2), [x1 | x2] = w12(n2)
                DAD_TRY(wgrad_linear(dy, Dm, bt.h, Hd, M, Dm, Hd, G(b + "mlp.w3.weight"), ar));
                DAD_TRY(bias_grad(dy, Dm, M, Dm, G(b + "mlp.w3.bias")));
                DAD_TRY(dgrad_linear(dy, Dm, M, Dm, m.P(b + "mlp.w3.weight"), Hd, dg, ar));
                RUN(swiglu_bwd(bt.hpre, dg, dh, bf, M, Hd, st));
                DAD_TRY(wgrad_linear(dh, 2 * Hd, bt.n2, Dm, M, 2 * Hd, Dm, G(b + "mlp.w12.weight"), ar));
                DAD_TRY(bias_grad(dh, 2 * Hd, M, 2 * Hd, G(b + "mlp.w12.bias")));
                DAD_TRY(dgrad_linear(dh, 2 * Hd, M, 2 * Hd, m.P(b + "mlp.w12.weight"), Dm, dn, ar));
            } else {
                DAD_TRY(wgrad_linear(dy, Dm, bt.h, 4 * Dm, M, Dm, 4 * Dm, G(b + "mlp.fc2.weight"), ar));
                DAD_TRY(bias_grad(dy, Dm, M, Dm, G(b + "mlp.fc2.bias")));
                DAD_TRY(dgrad_linear(dy, Dm, M, Dm, m.P(b + "mlp.fc2.weight"), 4 * Dm, dh, ar));
                RUN(gelu_bwd(bt.hpre, dh, dh, bf, M * 4 * Dm, st));
                DAD_TRY(wgrad_linear(dh, 4 * Dm, bt.n2, Dm, M, 4 * Dm, Dm, G(b + "mlp.fc1.weight"), ar));
                DAD_TRY(bias_grad(dh, 4 * Dm, M, 4 * Dm, G(b + "mlp.fc1.bias")));
                DAD_TRY(dgrad_linear(dh, 4 * Dm, M, 4 * Dm, m.P(b + "mlp.fc1.weight"), Dm, dn, ar));
            }
            RUN(layernorm_bwd(bt.x1, m.P(b + "norm2.weight"), dn, bf, Gx, G(b + "norm2.weight"), G(b + "norm2.bias"), M, Dm, 1, 1, 0,
                              LN_EPS, st));
            // x1 = x0 + gamma1 * y1
            RUN(colsum(Gx, 0, Dm, bt.y1, bf, Dm, M, Dm, G(b + "ls1.gamma"), m.P(b + "ls1.gamma"), dy, st));
            DAD_TRY(wgrad_linear(dy, Dm, bt.att, Dm, M, Dm, Dm, G(b + "attn.proj.weight"), ar));
            DAD_TRY(bias_grad(dy, Dm, M, Dm, G(b + "attn.proj.bias")));
            DAD_TRY(dgrad_linear(dy, Dm, M, Dm, m.P(b + "attn.proj.weight"), Dm, dn, ar));   // dn <- d att
            DAD_TRY(attention_bwd(bt, dn, dqkv, ar));
            DAD_TRY(wgrad_linear(dqkv, 3 * Dm, bt.n1, Dm, M, 3 * Dm, Dm, G(b + "attn.qkv.weight"), ar));
            DAD_TRY(bias_grad(dqkv, 3 * Dm, M, 3 * Dm, G(b + "attn.qkv.bias")));
            DAD_TRY(dgrad_linear(dqkv, 3 * Dm, M, 3 * Dm, m.P(b + "attn.qkv.weight"), Dm, dn, ar));
            RUN(layernorm_bwd(bt.x0, m.P(b + "norm1.weight"), dn, bf, Gx, G(b + "norm1.weight"), G(b + "norm1.bias"), M, Dm, 1, 1, 0,
                              LN_EPS, st));
            ar.used = mk;
        }
        // ---- patch embedding / positional table (fp32 engine: N = 588 columns, tiny)
        float* dWpe = G(p + "patch_embed.proj.weight");
        if (dWpe || dry) {   // (the dry run sizes the scratch for the all-parameters case)
            float* dW = dWpe;
            const size_t mk = ar.used;
            const float* ape_f = as_f32(t.ape, M * PATCH_KP, ar);
            if (!dry) {
                DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                SGemm g; g.A = Gx; g.sam = 1; g.sak = Dm; g.B = ape_f; g.sbk = PATCH_KP; g.sbn = 1; g.C = dW; g.scm = PATCH_K; g.scn = 1;
                g.M = Dm; g.N = PATCH_K; g.K = static_cast<int>(M); g.accumulate = 1;
                DAD_TRY(sgemm(g, st));
            }
            ar.used = mk;
        }
        float* dtab = ar.f(static_cast<size_t>(T) * Dm);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_TRY(batch_sum_rows(Gx, dtab, B, T, Dm, st));
            DAD_TRY(pos_table_bwd(dtab, G(p + "pos_embed"), G(p + "cls_token"), G(p + "patch_embed.proj.bias"), Dm, H, W, st));
        }
        return DAD_OK;
    }
#undef RUN
};
