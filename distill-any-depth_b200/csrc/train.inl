// Training forward (activations kept on a tape) and backward of the DepthAnythingV2 student in the fp32
// verification engine (reference: autograd over dpt.py:150-225, dinov2.py:212-321,
// util/blocks.py:29-148, driven by tools/train_distillation.py:1556-1575).  Included by model.cu inside namespace dad.
//
// One caller-owned workspace holds the tape (written by forward_train, read by backward) followed by the backward
// scratch; both calls lay it out with the same deterministic bump sequence, so no pointers are carried between them.
// Parameter gradients ACCUMULATE into the fp32 buffers registered with dad_model_set_grad (unregistered = frozen).

namespace {

struct Bump {  // bump allocator with release-to-mark and a high-water mark
    uint8_t* base;
    size_t cap;
    bool dry;
    size_t used = 0, peak = 0;
    bool overflow = false;
    Bump(void* b, size_t c, bool d) : base(reinterpret_cast<uint8_t*>(b)), cap(c), dry(d) {}
    float* f(size_t n) {
        used = (used + 1023) & ~size_t(1023);
        float* p = dry ? nullptr : reinterpret_cast<float*>(base + used);
        used += n * 4;
        if (used > peak) peak = used;
        if (!dry && used > cap) { overflow = true; p = nullptr; }
        return p;
    }
};

struct BlockTape { float *x0, *n1, *qkv, *att, *y1, *x1, *n2, *hpre, *h, *y2; };
struct FusionTape {
    float *t1a = nullptr, *s = nullptr, *sr = nullptr, *t1b = nullptr, *u = nullptr, *tmp = nullptr, *res = nullptr;
    const float *lat = nullptr, *lat_relu = nullptr, *path = nullptr;
    int H = 0, W = 0, Ho = 0, Wo = 0;
    bool has_path = false;
};
struct Tape {
    float* ape = nullptr;
    std::vector<BlockTape> blk;
    float* xfinal = nullptr;
    float *tap[4], *pj[4], *rj[4], *lrn[4], *lrn_relu[4];
    FusionTape fu[4];
    float *o1 = nullptr, *up = nullptr, *t32 = nullptr, *depth = nullptr;
    int hs[4], wsz[4];
};

}  // namespace

struct Trainer {
    Model& m;
    int B, H, W;
    bool dry;
    cudaStream_t st;
    int Dm, L, F, heads, ph, pw, np, T;
    long long M, Mp;
    const int* oc;

    Trainer(Model& model, int B_, int H_, int W_, bool dry_, cudaStream_t st_) : m(model), B(B_), H(H_), W(W_), dry(dry_), st(st_) {
        Dm = m.D(); L = m.desc.depth; F = m.desc.features; heads = m.desc.num_heads; oc = m.desc.out_channels;
        ph = H / 14; pw = W / 14; np = ph * pw; T = np + 1;
        M = static_cast<long long>(B) * T; Mp = static_cast<long long>(B) * np;
    }

#define RUN(expr) do { if (!dry) DAD_TRY(expr); } while (0)

    float* G(const std::string& name) const {
        auto it = m.grads.find(name);
        return it == m.grads.end() ? nullptr : it->second.first;
    }

    void plan(Bump& ar, Tape& t) const {
        t.ape = ar.f(M * PATCH_KP);
        t.blk.resize(L);
        for (int i = 0; i < L; ++i) {
            BlockTape& b = t.blk[i];
            b.x0 = ar.f(M * Dm); b.n1 = ar.f(M * Dm); b.qkv = ar.f(M * 3 * Dm); b.att = ar.f(M * Dm); b.y1 = ar.f(M * Dm);
            b.x1 = ar.f(M * Dm); b.n2 = ar.f(M * Dm); b.hpre = ar.f(M * 4 * Dm); b.h = ar.f(M * 4 * Dm); b.y2 = ar.f(M * Dm);
        }
        t.xfinal = ar.f(M * Dm);
        const int hs[4] = {4 * ph, 2 * ph, ph, (ph + 2 - 3) / 2 + 1};
        const int wz[4] = {4 * pw, 2 * pw, pw, (pw + 2 - 3) / 2 + 1};
        for (int j = 0; j < 4; ++j) {
            t.hs[j] = hs[j]; t.wsz[j] = wz[j];
            t.tap[j] = ar.f(Mp * Dm);
            t.pj[j] = ar.f(Mp * oc[j]);
            t.rj[j] = (j == 2) ? t.pj[j] : ar.f(static_cast<size_t>(B) * hs[j] * wz[j] * oc[j]);
            const size_t n = static_cast<size_t>(B) * hs[j] * wz[j] * F;
            t.lrn[j] = ar.f(n);
            t.lrn_relu[j] = ar.f(n);
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
            if (f.has_path) { f.t1a = ar.f(n); f.s = ar.f(n); f.sr = ar.f(n); }
            else { f.s = t.lrn[r]; f.sr = t.lrn_relu[r]; }
            f.t1b = ar.f(n); f.u = ar.f(n); f.tmp = ar.f(no); f.res = ar.f(no);
        }
        const int H1 = 2 * hs[0], W1 = 2 * wz[0], F2 = F / 2;
        t.o1 = ar.f(static_cast<size_t>(B) * H1 * W1 * F2);
        t.up = ar.f(static_cast<size_t>(B) * H * W * F2);
        t.t32 = ar.f(static_cast<size_t>(B) * H * W * 32);
        t.depth = ar.f(static_cast<size_t>(B) * H * W);
    }

    // ------------------------------------------------------------------------------------ forward
    int forward(const float* x, float* depth_out, float* feat_out, Bump& ar, Tape& t) {
        plan(ar, t);
        // transient im2col buffer of the stride-2 reassemble conv (lies in the backward-scratch region)
        const int Cp3 = cdiv(oc[3], 64) * 64;
        const long long rows3 = static_cast<long long>(B) * t.hs[3] * t.wsz[3];
        float* col = ar.f(rows3 * 9 * Cp3);
        if (dry) return DAD_OK;
        DAD_REQUIRE(!ar.overflow, "forward_train: workspace too small for the activation tape");
        const std::string p = "pretrained.";
        DAD_TRY(patch_im2col(x, t.ape, 0, B, H, W, PATCH_KP, st));
        {
            Epilogue e; e.rowtab = m.pos_tables[std::make_pair(H, W)]; e.rowtab_period = T; e.out = t.blk[0].x0;
            DAD_TRY(m.linear(1, t.ape, M, PATCH_KP, m.patch, e, false, st));
        }
        int tj = 0;
        for (int i = 0; i < L; ++i) {
            const std::string b = p + "blocks." + std::to_string(i) + ".";
            BlockTape& bt = t.blk[i];
            float* xnext = (i + 1 < L) ? t.blk[i + 1].x0 : t.xfinal;
            DAD_TRY(layernorm(bt.x0, m.P(b + "norm1.weight"), m.P(b + "norm1.bias"), bt.n1, 0, nullptr, M, Dm, 1, 1, 0, LN_EPS, st));
            Epilogue eq; eq.bias = m.bqkv_scaled + static_cast<long long>(i) * 3 * Dm; eq.out = bt.qkv;
            DAD_TRY(m.linear(1, bt.n1, M, Dm, m.qkv[i], eq, false, st));
            DAD_TRY(attention(bt.qkv, bt.att, 0, B, T, heads, st));
            Epilogue ep; ep.bias = m.P(b + "attn.proj.bias"); ep.out = bt.y1;
            DAD_TRY(m.linear(1, bt.att, M, Dm, m.proj[i], ep, false, st));
            DAD_TRY(ls_residual(bt.x0, bt.y1, m.P(b + "ls1.gamma"), bt.x1, M, Dm, st));
            DAD_TRY(layernorm(bt.x1, m.P(b + "norm2.weight"), m.P(b + "norm2.bias"), bt.n2, 0, nullptr, M, Dm, 1, 1, 0, LN_EPS, st));
            Epilogue e1; e1.bias = m.P(b + "mlp.fc1.bias"); e1.out = bt.hpre;
            DAD_TRY(m.linear(1, bt.n2, M, Dm, m.fc1[i], e1, false, st));
            DAD_TRY(gelu_fwd(bt.hpre, bt.h, M * 4 * Dm, st));
            Epilogue e2; e2.bias = m.P(b + "mlp.fc2.bias"); e2.out = bt.y2;
            DAD_TRY(m.linear(1, bt.h, M, 4 * Dm, m.fc2[i], e2, false, st));
            DAD_TRY(ls_residual(bt.x1, bt.y2, m.P(b + "ls2.gamma"), xnext, M, Dm, st));
            if (tj < 4 && i == m.desc.taps[tj]) {
                DAD_TRY(layernorm(xnext, m.P(p + "norm.weight"), m.P(p + "norm.bias"), t.tap[tj], 0, (tj == 3) ? feat_out : nullptr,
                                  Mp, Dm, np, T, 1, LN_EPS, st));
                ++tj;
            }
        }
        DAD_REQUIRE(tj == 4, "taps must be increasing block indices < depth");

        const std::string h = "depth_head.", s = h + "scratch.";
        for (int j = 0; j < 4; ++j) {
            Epilogue e; e.bias = m.P(h + "projects." + std::to_string(j) + ".bias"); e.out = t.pj[j];
            DAD_TRY(m.linear(1, t.tap[j], Mp, Dm, m.projects[j], e, false, st));
            if (j == 0 || j == 1) {
                Epilogue es_; es_.bias = m.P(h + "resize_layers." + std::to_string(j) + ".bias"); es_.out = t.rj[j];
                es_.ldc = oc[j]; es_.scat_k = j == 0 ? 4 : 2; es_.scat_CoP = j == 0 ? m.CoP0 : m.CoP1; es_.scat_Co = oc[j];
                es_.scat_H = ph; es_.scat_W = pw;
                DAD_TRY(m.linear(1, t.pj[j], Mp, oc[j], j == 0 ? m.resize0 : m.resize1, es_, false, st));
            } else if (j == 3) {
                Epilogue e3; e3.bias = m.P(h + "resize_layers.3.bias"); e3.out = t.rj[3];
                DAD_TRY(im2col_s2(t.pj[3], col, 0, B, ph, pw, oc[3], Cp3, st));
                DAD_TRY(m.linear(1, col, rows3, 9 * Cp3, m.resize3, e3, false, st));
            }
            Epilogue er; er.out = t.lrn[j]; er.out_relu = t.lrn_relu[j];
            DAD_TRY(m.conv(1, t.rj[j], B, t.hs[j], t.wsz[j], oc[j], m.layer_rn[j], 9, er, false, st));
        }
        for (int r = 3; r >= 0; --r) {
            FusionTape& f = t.fu[r];
            const std::string q = s + "refinenet" + std::to_string(r + 1) + ".";
            if (f.has_path) {
                Epilogue e1; e1.bias = m.P(q + "resConfUnit1.conv1.bias"); e1.act = ACT_RELU; e1.out = f.t1a;
                DAD_TRY(m.conv(1, f.lat_relu, B, f.H, f.W, F, m.rcu[r][0][0], 9, e1, false, st));
                Epilogue e2; e2.bias = m.P(q + "resConfUnit1.conv2.bias"); e2.res1 = f.lat; e2.res2 = f.path; e2.out = f.s;
                e2.out_relu = f.sr;
                DAD_TRY(m.conv(1, f.t1a, B, f.H, f.W, F, m.rcu[r][0][1], 9, e2, false, st));
            }
            Epilogue e1; e1.bias = m.P(q + "resConfUnit2.conv1.bias"); e1.act = ACT_RELU; e1.out = f.t1b;
            DAD_TRY(m.conv(1, f.sr, B, f.H, f.W, F, m.rcu[r][1][0], 9, e1, false, st));
            Epilogue e2; e2.bias = m.P(q + "resConfUnit2.conv2.bias"); e2.res1 = f.s; e2.out = f.u;
            DAD_TRY(m.conv(1, f.t1b, B, f.H, f.W, F, m.rcu[r][1][1], 9, e2, false, st));
            DAD_TRY(bilinear_nhwc(f.u, f.tmp, 0, B, f.H, f.W, f.Ho, f.Wo, F, st));
            Epilogue eo; eo.bias = m.P(q + "out_conv.bias"); eo.out = f.res;
            DAD_TRY(m.conv(1, f.tmp, B, f.Ho, f.Wo, F, m.out_conv[r], 1, eo, false, st));
        }
        const int H1 = 2 * t.hs[0], W1 = 2 * t.wsz[0], F2 = F / 2;
        Epilogue eo1; eo1.bias = m.P(s + "output_conv1.bias"); eo1.out = t.o1;
        DAD_TRY(m.conv(1, t.fu[0].res, B, H1, W1, F, m.output_conv1, 9, eo1, false, st));
        DAD_TRY(bilinear_nhwc(t.o1, t.up, 0, B, H1, W1, H, W, F2, st));
        Epilogue eh; eh.bias = m.P(s + "output_conv2.0.bias"); eh.act = ACT_RELU; eh.out = t.t32;
        DAD_TRY(m.conv(1, t.up, B, H, W, F2, m.output_conv2_0, 9, eh, false, st));
        const long long P = static_cast<long long>(B) * H * W;
        DAD_TRY(head1x1(t.t32, m.P(s + "output_conv2.2.weight"), m.head_bias_host, t.depth, P, st));
        DAD_CHECK_CUDA(cudaMemcpyAsync(depth_out, t.depth, P * 4, cudaMemcpyDeviceToDevice, st));
        return DAD_OK;
    }

    // ------------------------------------------------------------------------------------ backward helpers
    // dW[Nout, Kin] += dY^T X
    int wgrad_linear(const float* dY, long long ldy, const float* X, long long ldx, long long rows, int Nout, int Kin, float* dW) {
        if (!dW || dry) return DAD_OK;
        SGemm g; g.A = dY; g.sam = 1; g.sak = ldy; g.B = X; g.sbk = ldx; g.sbn = 1; g.C = dW; g.scm = Kin; g.scn = 1;
        g.M = Nout; g.N = Kin; g.K = static_cast<int>(rows); g.accumulate = 1;
        return sgemm(g, st);
    }
    // dX[rows, Kin] = dY[rows, Nout] W[Nout, Kin]
    int dgrad_linear(const float* dY, long long ldy, long long rows, int Nout, const float* Wm, int Kin, float* dX) {
        if (dry) return DAD_OK;
        SGemm g; g.A = dY; g.sam = ldy; g.sak = 1; g.B = Wm; g.sbk = Kin; g.sbn = 1; g.C = dX; g.scm = Kin; g.scn = 1;
        g.M = static_cast<int>(rows); g.N = Kin; g.K = Nout;
        return sgemm(g, st);
    }
    int bias_grad(const float* dY, long long ld, long long rows, int N, float* db) {
        if (!db || dry) return DAD_OK;
        return colsum(dY, ld, nullptr, 0, rows, N, db, nullptr, nullptr, st);
    }
    // dW[Co, Ci, taps] += sum_pixels dOut[p, co] * window(X)[p, tap, ci];  X is [B, Hin, Win, Ci], dOut [B, Ho, Wo, Co]
    int conv_wgrad(const float* X, const float* dOut, int Hin, int Win, int Ci, int Co, int taps, int stride, int Ho, int Wo,
                   float* dW) {
        if (!dW || dry) return DAD_OK;
        SGemm g; g.A = dOut; g.sam = 1; g.sak = Co; g.B = X; g.C = dW;
        g.M = Co; g.N = taps * Ci; g.K = B * Ho * Wo; g.accumulate = 1;
        g.conv_taps = taps; g.convC = Ci; g.convH = Hin; g.convW = Win; g.convHo = Ho; g.convWo = Wo; g.conv_stride = stride;
        g.cmap = 1;
        return sgemm(g, st);
    }
    // stride-1 conv data gradient through the forward conv engine with flipped / transposed weights:
    // dIn[B,Hc,Wc,Ci] = (add ? add : 0) + conv(dOut[B,Hc,Wc,Co], Wd)
    int conv_dgrad(const float* dOut, int Hc, int Wc, int Co, int Ci, int taps, const float* Wmaster, float* dIn, const float* add,
                   Bump& ar) {
        const int CoP = cdiv(Co, 64) * 64;
        const size_t mk = ar.used;
        float* wd = ar.f(static_cast<size_t>(Ci) * taps * CoP);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_TRY(pack_conv_dgrad(Wmaster, wd, Co, Ci, taps, CoP, st));
            Mat mt; mt.w[1] = wd; mt.N = Ci; mt.Kp = taps * CoP;
            Epilogue e; e.out = dIn;
            if (add) e.res1 = add;
            DAD_TRY(m.conv(1, dOut, B, Hc, Wc, Co, mt, taps, e, false, st));
        }
        ar.used = mk;  // stream order keeps wd alive until the conv has read it; the next user writes after it
        return DAD_OK;
    }

    int attention_bwd(const BlockTape& bt, const float* datt, float* dqkv, Bump& ar) {
        const size_t mk = ar.used;
        const long long tt = static_cast<long long>(T) * T;
        float* Pm = ar.f(static_cast<size_t>(B) * heads * tt);
        float* dP = ar.f(static_cast<size_t>(B) * heads * tt);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            const long long ld = 3LL * Dm;
            SGemm s;  // S = Q' K^T
            s.A = bt.qkv; s.sam = ld; s.sak = 1; s.a1 = T * ld; s.a2 = 64;
            s.B = bt.qkv + Dm; s.sbk = 1; s.sbn = ld; s.b1 = T * ld; s.b2 = 64;
            s.C = Pm; s.scm = T; s.scn = 1; s.c1 = heads * tt; s.c2 = tt;
            s.M = T; s.N = T; s.K = 64; s.nb1 = B; s.nb2 = heads;
            DAD_TRY(sgemm(s, st));
            DAD_TRY(softmax_rows(Pm, static_cast<long long>(B) * heads * T, T, st));
            SGemm p;  // dP = dO V^T
            p.A = datt; p.sam = Dm; p.sak = 1; p.a1 = static_cast<long long>(T) * Dm; p.a2 = 64;
            p.B = bt.qkv + 2 * Dm; p.sbk = 1; p.sbn = ld; p.b1 = T * ld; p.b2 = 64;
            p.C = dP; p.scm = T; p.scn = 1; p.c1 = heads * tt; p.c2 = tt;
            p.M = T; p.N = T; p.K = 64; p.nb1 = B; p.nb2 = heads;
            DAD_TRY(sgemm(p, st));
            SGemm v;  // dV = P^T dO
            v.A = Pm; v.sam = 1; v.sak = T; v.a1 = heads * tt; v.a2 = tt;
            v.B = datt; v.sbk = Dm; v.sbn = 1; v.b1 = static_cast<long long>(T) * Dm; v.b2 = 64;
            v.C = dqkv + 2 * Dm; v.scm = ld; v.scn = 1; v.c1 = T * ld; v.c2 = 64;
            v.M = T; v.N = 64; v.K = T; v.nb1 = B; v.nb2 = heads;
            DAD_TRY(sgemm(v, st));
            DAD_TRY(softmax_bwd_rows(Pm, dP, static_cast<long long>(B) * heads * T, T, st));   // dP <- dS
            SGemm q;  // dq = 0.125 * dS K   (the packed q rows carry 64^-0.5: q' = q / 8)
            q.A = dP; q.sam = T; q.sak = 1; q.a1 = heads * tt; q.a2 = tt;
            q.B = bt.qkv + Dm; q.sbk = ld; q.sbn = 1; q.b1 = T * ld; q.b2 = 64;
            q.C = dqkv; q.scm = ld; q.scn = 1; q.c1 = T * ld; q.c2 = 64;
            q.M = T; q.N = 64; q.K = T; q.nb1 = B; q.nb2 = heads; q.alpha = 0.125f;
            DAD_TRY(sgemm(q, st));
            SGemm k;  // dK = dS^T Q'
            k.A = dP; k.sam = 1; k.sak = T; k.a1 = heads * tt; k.a2 = tt;
            k.B = bt.qkv; k.sbk = ld; k.sbn = 1; k.b1 = T * ld; k.b2 = 64;
            k.C = dqkv + Dm; k.scm = ld; k.scn = 1; k.c1 = T * ld; k.c2 = 64;
            k.M = T; k.N = 64; k.K = T; k.nb1 = B; k.nb2 = heads;
            DAD_TRY(sgemm(k, st));
        }
        ar.used = mk;
        return DAD_OK;
    }

    // one FeatureFusionBlock backward (util/blocks.py:129-146); dres [B,Ho,Wo,F] -> dlat [B,H,W,F], dpath (has_path)
    int fusion_bwd(int r, const FusionTape& f, const float* dres, float** dlat_out, float** dpath_out, Bump& ar) {
        const std::string q = "depth_head.scratch.refinenet" + std::to_string(r + 1) + ".";
        const long long n = static_cast<long long>(B) * f.H * f.W * F, no = static_cast<long long>(B) * f.Ho * f.Wo * F;
        const long long px = n / F, pxo = no / F;
        DAD_TRY(conv_wgrad(f.tmp, dres, f.Ho, f.Wo, F, F, 1, 1, f.Ho, f.Wo, G(q + "out_conv.weight")));
        DAD_TRY(bias_grad(dres, F, pxo, F, G(q + "out_conv.bias")));
        float* dtmp = ar.f(no);
        DAD_TRY(conv_dgrad(dres, f.Ho, f.Wo, F, F, 1, m.P(q + "out_conv.weight"), dtmp, nullptr, ar));
        float* du = ar.f(n);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_CHECK_CUDA(cudaMemsetAsync(du, 0, n * 4, st));
            DAD_TRY(bilinear_bwd(dtmp, du, B, f.H, f.W, f.Ho, f.Wo, F, st));
        }
        // RCU2: u = conv2(relu(conv1(relu(s)) + b1)) + b2 + s
        DAD_TRY(conv_wgrad(f.t1b, du, f.H, f.W, F, F, 9, 1, f.H, f.W, G(q + "resConfUnit2.conv2.weight")));
        DAD_TRY(bias_grad(du, F, px, F, G(q + "resConfUnit2.conv2.bias")));
        float* dt1b = ar.f(n);
        DAD_TRY(conv_dgrad(du, f.H, f.W, F, F, 9, m.P(q + "resConfUnit2.conv2.weight"), dt1b, nullptr, ar));
        RUN(relu_bwd(dt1b, f.t1b, nullptr, dt1b, n, st));
        DAD_TRY(conv_wgrad(f.sr, dt1b, f.H, f.W, F, F, 9, 1, f.H, f.W, G(q + "resConfUnit2.conv1.weight")));
        DAD_TRY(bias_grad(dt1b, F, px, F, G(q + "resConfUnit2.conv1.bias")));
        float* ds = ar.f(n);
        DAD_TRY(conv_dgrad(dt1b, f.H, f.W, F, F, 9, m.P(q + "resConfUnit2.conv1.weight"), ds, nullptr, ar));
        RUN(relu_bwd(ds, f.sr, du, ds, n, st));   // ds = du + dsr * (s > 0)
        if (!f.has_path) {
            *dlat_out = ds;
            *dpath_out = nullptr;
            return DAD_OK;
        }
        // RCU1 on the lateral: s = conv2(relu(conv1(relu(lat)) + b1)) + b2 + lat + path
        DAD_TRY(conv_wgrad(f.t1a, ds, f.H, f.W, F, F, 9, 1, f.H, f.W, G(q + "resConfUnit1.conv2.weight")));
        DAD_TRY(bias_grad(ds, F, px, F, G(q + "resConfUnit1.conv2.bias")));
        float* dt1a = ar.f(n);
        DAD_TRY(conv_dgrad(ds, f.H, f.W, F, F, 9, m.P(q + "resConfUnit1.conv2.weight"), dt1a, nullptr, ar));
        RUN(relu_bwd(dt1a, f.t1a, nullptr, dt1a, n, st));
        DAD_TRY(conv_wgrad(f.lat_relu, dt1a, f.H, f.W, F, F, 9, 1, f.H, f.W, G(q + "resConfUnit1.conv1.weight")));
        DAD_TRY(bias_grad(dt1a, F, px, F, G(q + "resConfUnit1.conv1.bias")));
        float* dlat = ar.f(n);
        DAD_TRY(conv_dgrad(dt1a, f.H, f.W, F, F, 9, m.P(q + "resConfUnit1.conv1.weight"), dlat, nullptr, ar));
        RUN(relu_bwd(dlat, f.lat_relu, ds, dlat, n, st));   // dlat = ds + dlr * (lat > 0)
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

        // ---- output head
        float* dt32 = ar.f(P * 32);
        if (!dry) DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
        RUN(head_bwd(gdepth, t.depth, t.t32, m.P(s + "output_conv2.2.weight"), dt32, G(s + "output_conv2.2.weight"),
                     G(s + "output_conv2.2.bias"), P, st));
        DAD_TRY(conv_wgrad(t.up, dt32, H, W, F2, 32, 9, 1, H, W, G(s + "output_conv2.0.weight")));
        DAD_TRY(bias_grad(dt32, 32, P, 32, G(s + "output_conv2.0.bias")));
        float* dup = ar.f(P * F2);
        DAD_TRY(conv_dgrad(dt32, H, W, 32, F2, 9, m.P(s + "output_conv2.0.weight"), dup, nullptr, ar));
        float* do1 = ar.f(P1 * F2);
        if (!dry) {
            DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_CHECK_CUDA(cudaMemsetAsync(do1, 0, P1 * F2 * 4, st));
            DAD_TRY(bilinear_bwd(dup, do1, B, H1, W1, H, W, F2, st));
        }
        DAD_TRY(conv_wgrad(t.fu[0].res, do1, H1, W1, F, F2, 9, 1, H1, W1, G(s + "output_conv1.weight")));
        DAD_TRY(bias_grad(do1, F2, P1, F2, G(s + "output_conv1.bias")));
        float* dres = ar.f(P1 * F);
        DAD_TRY(conv_dgrad(do1, H1, W1, F2, F, 9, m.P(s + "output_conv1.weight"), dres, nullptr, ar));

        // ---- fusion blocks, finest first
        float* dlat[4];
        for (int r = 0; r < 4; ++r) {
            float* dpath = nullptr;
            DAD_TRY(fusion_bwd(r, t.fu[r], dres, &dlat[r], &dpath, ar));
            dres = dpath;
        }

        // ---- reassemble: layer_rn -> resize -> projects; dtap[j] = gradient of the LayerNorm'd tap
        float* dtap[4];
        for (int j = 0; j < 4; ++j) {
            const std::string js = std::to_string(j);
            const int hj = t.hs[j], wj = t.wsz[j];
            const long long pxj = static_cast<long long>(B) * hj * wj;
            DAD_TRY(conv_wgrad(t.rj[j], dlat[j], hj, wj, oc[j], F, 9, 1, hj, wj, G(s + "layer" + std::to_string(j + 1) + "_rn.weight")));
            float* drj = ar.f(pxj * oc[j]);
            DAD_TRY(conv_dgrad(dlat[j], hj, wj, F, oc[j], 9, m.P(s + "layer" + std::to_string(j + 1) + "_rn.weight"), drj, nullptr, ar));
            float* dpj = drj;
            if (j == 0 || j == 1) {
                const int k = j == 0 ? 4 : 2, kk = k * k, CoP = j == 0 ? m.CoP0 : m.CoP1;
                const Mat& mt = j == 0 ? m.resize0 : m.resize1;
                DAD_TRY(bias_grad(drj, oc[j], pxj, oc[j], G(h + "resize_layers." + js + ".bias")));
                float* Gm = ar.f(Mp * kk * CoP);
                dpj = ar.f(Mp * oc[j]);
                if (!dry) {
                    DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                    DAD_TRY(convT_gather(drj, Gm, B, ph, pw, k, oc[j], CoP, st));
                    if (float* dW = G(h + "resize_layers." + js + ".weight")) {
                        SGemm g; g.A = Gm; g.sam = 1; g.sak = static_cast<long long>(kk) * CoP; g.B = t.pj[j]; g.sbk = oc[j]; g.sbn = 1;
                        g.C = dW; g.M = kk * CoP; g.N = oc[j]; g.K = static_cast<int>(Mp); g.accumulate = 1;
                        g.cmap = 2; g.ct_CoP = CoP; g.ct_Co = oc[j]; g.ct_kk = kk;
                        DAD_TRY(sgemm(g, st));
                    }
                    // dpj[Mp, Ci] = Gm[Mp, kk*CoP] Wm[kk*CoP, Ci]  (the packed fp32 ConvTranspose matrix)
                    SGemm d; d.A = Gm; d.sam = static_cast<long long>(kk) * CoP; d.sak = 1; d.B = reinterpret_cast<const float*>(mt.w[1]);
                    d.sbk = mt.Kp; d.sbn = 1; d.C = dpj; d.scm = oc[j]; d.scn = 1; d.M = static_cast<int>(Mp); d.N = oc[j]; d.K = kk * CoP;
                    DAD_TRY(sgemm(d, st));
                }
            } else if (j == 3) {
                const int Cp = cdiv(oc[3], 64) * 64;
                DAD_TRY(bias_grad(drj, oc[3], pxj, oc[3], G(h + "resize_layers.3.bias")));
                DAD_TRY(conv_wgrad(t.pj[3], drj, ph, pw, oc[3], oc[3], 9, 2, hj, wj, G(h + "resize_layers.3.weight")));
                float* dcol = ar.f(pxj * 9 * Cp);
                dpj = ar.f(Mp * oc[3]);
                if (!dry) {
                    DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
                    SGemm d; d.A = drj; d.sam = oc[3]; d.sak = 1; d.B = reinterpret_cast<const float*>(m.resize3.w[1]); d.sbk = 9 * Cp;
                    d.sbn = 1; d.C = dcol; d.scm = 9 * Cp; d.scn = 1; d.M = static_cast<int>(pxj); d.N = 9 * Cp; d.K = oc[3];
                    DAD_TRY(sgemm(d, st));
                    DAD_TRY(col2im_s2(dcol, dpj, B, ph, pw, oc[3], Cp, st));
                }
            }
            DAD_TRY(wgrad_linear(dpj, oc[j], t.tap[j], Dm, Mp, oc[j], Dm, G(h + "projects." + js + ".weight")));
            DAD_TRY(bias_grad(dpj, oc[j], Mp, oc[j], G(h + "projects." + js + ".bias")));
            dtap[j] = ar.f(Mp * Dm);
            if (!dry) DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            DAD_TRY(dgrad_linear(dpj, oc[j], Mp, oc[j], m.P(h + "projects." + js + ".weight"), Dm, dtap[j]));
        }
        if (gfeat) RUN(add_inplace(dtap[3], gfeat, Mp * Dm, st));

        // ---- encoder, last block first.  Gx = gradient of the residual stream
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
                RUN(layernorm_bwd(xnext, m.P(p + "norm.weight"), dtap[tj], Gx, G(p + "norm.weight"), G(p + "norm.bias"), Mp, Dm, np, T,
                                  1, LN_EPS, st));
                --tj;
            }
            const size_t mk = ar.used;
            float* dy = ar.f(M * Dm);        // gradient of the branch output before LayerScale
            float* dh = ar.f(M * 4 * Dm);
            float* dn = ar.f(M * Dm);
            float* dqkv = ar.f(M * 3 * Dm);
            if (!dry) DAD_REQUIRE(!ar.overflow, "backward: workspace too small");
            // x_{i+1} = x1 + gamma2 * y2
            RUN(colsum(Gx, Dm, bt.y2, Dm, M, Dm, G(b + "ls2.gamma"), m.P(b + "ls2.gamma"), dy, st));
            DAD_TRY(wgrad_linear(dy, Dm, bt.h, 4 * Dm, M, Dm, 4 * Dm, G(b + "mlp.fc2.weight")));
            DAD_TRY(bias_grad(dy, Dm, M, Dm, G(b + "mlp.fc2.bias")));
            DAD_TRY(dgrad_linear(dy, Dm, M, Dm, m.P(b + "mlp.fc2.weight"), 4 * Dm, dh));
            RUN(gelu_bwd(bt.hpre, dh, dh, M * 4 * Dm, st));
            DAD_TRY(wgrad_linear(dh, 4 * Dm, bt.n2, Dm, M, 4 * Dm, Dm, G(b + "mlp.fc1.weight")));
            DAD_TRY(bias_grad(dh, 4 * Dm, M, 4 * Dm, G(b + "mlp.fc1.bias")));
            DAD_TRY(dgrad_linear(dh, 4 * Dm, M, 4 * Dm, m.P(b + "mlp.fc1.weight"), Dm, dn));
            RUN(layernorm_bwd(bt.x1, m.P(b + "norm2.weight"), dn, Gx, G(b + "norm2.weight"), G(b + "norm2.bias"), M, Dm, 1, 1, 0,
                              LN_EPS, st));
            // x1 = x0 + gamma1 * y1
            RUN(colsum(Gx, Dm, bt.y1, Dm, M, Dm, G(b + "ls1.gamma"), m.P(b + "ls1.gamma"), dy, st));
            DAD_TRY(wgrad_linear(dy, Dm, bt.att, Dm, M, Dm, Dm, G(b + "attn.proj.weight")));
            DAD_TRY(bias_grad(dy, Dm, M, Dm, G(b + "attn.proj.bias")));
            DAD_TRY(dgrad_linear(dy, Dm, M, Dm, m.P(b + "attn.proj.weight"), Dm, dn));   // dn <- d att
            DAD_TRY(attention_bwd(bt, dn, dqkv, ar));
            DAD_TRY(wgrad_linear(dqkv, 3 * Dm, bt.n1, Dm, M, 3 * Dm, Dm, G(b + "attn.qkv.weight")));
            DAD_TRY(bias_grad(dqkv, 3 * Dm, M, 3 * Dm, G(b + "attn.qkv.bias")));
            DAD_TRY(dgrad_linear(dqkv, 3 * Dm, M, 3 * Dm, m.P(b + "attn.qkv.weight"), Dm, dn));
            RUN(layernorm_bwd(bt.x0, m.P(b + "norm1.weight"), dn, Gx, G(b + "norm1.weight"), G(b + "norm1.bias"), M, Dm, 1, 1, 0,
                              LN_EPS, st));
            ar.used = mk;
        }
        // ---- patch embedding / positional table
        if (float* dW = G(p + "patch_embed.proj.weight")) {
            if (!dry) {
                SGemm g; g.A = Gx; g.sam = 1; g.sak = Dm; g.B = t.ape; g.sbk = PATCH_KP; g.sbn = 1; g.C = dW; g.scm = PATCH_K; g.scn = 1;
                g.M = Dm; g.N = PATCH_K; g.K = static_cast<int>(M); g.accumulate = 1;
                DAD_TRY(sgemm(g, st));
            }
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
