// Inception-v3 trunk: weight packing, workspace plan and forward (host side).
// Topology = torchvision.models.Inception3 as wired by image_realism/FID/inception.py:59-95
// (block0: Conv2d_1a/2a/2b + maxpool, block1: Conv2d_3b/4a + maxpool, block2: Mixed_5b..6e,
//  block3: Mixed_7a..7c + adaptive avg pool).  Branch outputs are written straight into channel
// slices of the block output, in torch.cat order.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "net.cuh"

namespace tise {

int fail(tise_ctx* ctx, int status, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return status;
}

int conv_out_dim(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }

static int dev_alloc(tise_ctx* ctx, std::vector<void*>* allocs, void** p, size_t bytes) {
  TISE_CUDA(ctx, cudaMalloc(p, bytes));
  allocs->push_back(*p);
  return TISE_OK;
}

// ---------------------------------------------------------------------------------------------
// Packing: w [cout][cin][R][S] fp32 -> fp16 hi/lo rows in the layout described at ConvLayer (common.cuh).  Each
// output channel is scaled by 2^e so that its largest |w| lands in [2^11, 2^12): the lo plane then stays in
// fp16's normal range, and 2^-e is folded into the epilogue scale (exact).
// ---------------------------------------------------------------------------------------------
int conv_chunk_of(int cin) {
  // TMA cost was measured to be per row (one pixel's / one output channel's chunk), ~2 cycles regardless of the
  // row width, so narrower chunks only add rows: keep 64-channel chunks and let TMA zero-fill the remainder
  // (the kernel supports kc = 32 and 16 -- SWIZZLE_64B / 32B operand tiles -- via TISE_UMMA_KC for experiments).
  const char* e = getenv("TISE_UMMA_KC");
  if (e && atoi(e) == 1) {
    if (cin % 64 == 0) return 64;
    if (cin % 32 == 0) return 32;
    if (cin == 48) return 64;
    return 16;
  }
  (void)cin;
  return 64;
}
// Largest divisor of cout that is a multiple of 16 and at most 256.
// N tile of a layer: a divisor of cout, multiple of 16, <= 256 (one tcgen05.mma).
// * CTA pairs (default): the largest one -- per K block a CTA stages 16 KiB of activations plus nt x 128 B of weights
//   and re-reads them once per K step, so shared-memory bytes per tensor-pipe cycle fall monotonically with nt.
// * single-CTA form (TISE_UMMA_PAIR=0): the divisor that needs the fewest fill bytes per cycle when tiles of
//   nt <= 128 pair two M tiles per work item (conv_umma.cu).
int conv_ntile_of(int cout) {
  if (umma_pair_mode()) {
    for (int nt = 256; nt >= 16; nt -= 16)
      if (cout % nt == 0) return nt;
    return 0;
  }
  static int pair_m = -1;
  if (pair_m < 0) pair_m = (getenv("TISE_UMMA_MSUB") && atoi(getenv("TISE_UMMA_MSUB")) == 1) ? 0 : 1;
  int best = 0;
  double best_cost = 1e30;
  for (int nt = 256; nt >= 16; nt -= 16) {
    if (cout % nt) continue;
    const int msub = (pair_m && nt <= 128) ? 2 : 1;
    const double cost = (msub * 16384.0 + 256.0 * nt) / (msub * 4.0 * nt);
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = nt;
    }
  }
  return best;
}

int pack_conv_layer(tise_ctx* ctx, ConvLayer* L, const float* w, const float* scale, const float* shift,
                    std::vector<void*>* allocs, int force_kc) {
  L->kc = force_kc ? force_kc : conv_chunk_of(L->cin);
  L->cin_pad = (L->cin + L->kc - 1) / L->kc * L->kc;
  L->ktot = L->R * L->S * L->cin_pad;
  L->nt = conv_ntile_of(L->cout);
  if (L->cout % 16 || L->cin % 16 || L->nt == 0)
    return fail(ctx, TISE_ERR_INVALID, "conv layer %d -> %d channels: both must be multiples of 16", L->cin, L->cout);
  const size_t n = (size_t)2 * L->cout * L->ktot;
  std::vector<__half> wp(n, __float2half_rn(0.f));
  std::vector<float> sc(L->cout), sh(L->cout);
  const int taps = L->R * L->S;
  for (int co = 0; co < L->cout; ++co) {
    const float* wc = w + (size_t)co * L->cin * taps;
    float maxabs = 0.f;
    for (int i = 0; i < L->cin * taps; ++i) maxabs = fmaxf(maxabs, fabsf(wc[i]));
    int e = 0;
    if (maxabs > 0.f && isfinite(maxabs)) e = 11 - ilogbf(maxabs);
    __half* hi = wp.data() + conv_weight_row(*L, co, 0) * L->ktot;
    __half* lo = wp.data() + conv_weight_row(*L, co, 1) * L->ktot;
    for (int ci = 0; ci < L->cin; ++ci)
      for (int t = 0; t < taps; ++t) {
        const float v = ldexpf(wc[(size_t)ci * taps + t], e);
        const __half h = __float2half_rn(v);
        const size_t k = (size_t)t * L->cin_pad + ci;
        hi[k] = h;
        lo[k] = __float2half_rn(v - __half2float(h));
      }
    sc[co] = ldexpf(scale[co], -e);
    sh[co] = shift[co];
  }
  int rc;
  if ((rc = dev_alloc(ctx, allocs, (void**)&L->d_w, n * sizeof(__half)))) return rc;
  if ((rc = dev_alloc(ctx, allocs, (void**)&L->d_scale, L->cout * sizeof(float)))) return rc;
  if ((rc = dev_alloc(ctx, allocs, (void**)&L->d_shift, L->cout * sizeof(float)))) return rc;
  if ((rc = dev_alloc(ctx, allocs, (void**)&L->d_shift1, L->cout * sizeof(float)))) return rc;
  TISE_CUDA(ctx, cudaMemcpy(L->d_w, wp.data(), n * sizeof(__half), cudaMemcpyHostToDevice));
  TISE_CUDA(ctx, cudaMemcpy(L->d_scale, sc.data(), L->cout * sizeof(float), cudaMemcpyHostToDevice));
  TISE_CUDA(ctx, cudaMemcpy(L->d_shift, sh.data(), L->cout * sizeof(float), cudaMemcpyHostToDevice));
  TISE_CUDA(ctx, cudaMemcpy(L->d_shift1, sh.data(), L->cout * sizeof(float), cudaMemcpyHostToDevice));
  return TISE_OK;
}

// eval-mode BatchNorm2d(eps=1e-3) of torchvision BasicConv2d -> per-channel scale / shift
static int fold_bn(tise_ctx* ctx, Net* net, const std::string& prefix, int cout, std::vector<float>* scale,
                   std::vector<float>* shift) {
  const char* names[4] = {".bn.weight", ".bn.bias", ".bn.running_mean", ".bn.running_var"};
  const std::vector<float>* t[4];
  for (int i = 0; i < 4; ++i) {
    auto it = net->raw.find(prefix + names[i]);
    if (it == net->raw.end() || (int)it->second.size() != cout)
      return fail(ctx, TISE_ERR_INVALID, "state_dict: missing or mis-sized %s%s", prefix.c_str(), names[i]);
    t[i] = &it->second;
  }
  scale->resize(cout);
  shift->resize(cout);
  for (int c = 0; c < cout; ++c) {
    const double g = (*t[0])[c], b = (*t[1])[c], m = (*t[2])[c], v = (*t[3])[c];
    const double s = g / sqrt(v + 1e-3);
    (*scale)[c] = (float)s;
    (*shift)[c] = (float)(b - m * s);
  }
  return TISE_OK;
}

// ---------------------------------------------------------------------------------------------
// Builder: walks the topology once; creates layers on first sight (net_load) and emits steps.
// ---------------------------------------------------------------------------------------------
struct Buf {
  size_t off;  // per-image byte offset; the buffer of a batch of n starts at ws + off * n
  int h, w, c, f32;
};
struct View {
  int buf, choff, c;
};

struct Builder {
  tise_ctx* ctx;
  Net* net;
  Plan* plan;  // may be null (declare-only walk)
  int n;
  char* ws;
  bool flat_stem = true;  // tcgen05 path; the SIMT cross-check path keeps dense 147 x 147 tensors
  std::vector<Buf> bufs;
  size_t per_image = 0;
  int rc = TISE_OK;

  int buf(int h, int w, int c, int f32 = 0) {
    Buf b{per_image, h, w, c, f32};
    per_image += align_up((size_t)h * w * c * (f32 ? 4 : 2), 256);
    bufs.push_back(b);
    return (int)bufs.size() - 1;
  }
  View all(int b) { return View{b, 0, bufs[b].c}; }
  char* ptr(const View& v) {
    const Buf& b = bufs[v.buf];
    return ws + b.off * (size_t)n + (size_t)v.choff * (b.f32 ? 4 : 2);
  }

  // One member of a (possibly fused) convolution: a torchvision BasicConv2d by name, its destination view,
  // and whether it is written raw (branch_pool: BN + ReLU are applied after the average pooling).
  struct Member {
    std::string name;
    View out;
    bool raw;
  };

  // Creates (once) the packed layer holding the members' output channels back to back.
  // fold_s: the S taps of a filter row become channels (cin' = S * cin, S' = 1): for stride-1, unpadded layers a
  // window of S neighbouring NHWC pixels IS one contiguous run of S * cin channels, so e.g. Conv2d_4a (cin = 80:
  // 9 taps x 2 chunks of which the second holds 16 channels) walks 3 x 4 K blocks instead of 18.
  const ConvLayer* layer(const std::string& key, const std::vector<Member>& ms, int cin, int R, int S, int stride, int ph,
                         int pw, int force_kc = 0, bool fold_s = false) {
    auto it = net->convs.find(key);
    if (it != net->convs.end()) return &it->second;
    const int taps = R * S;
    int cout = 0;
    for (const Member& m : ms) cout += m.out.c;
    std::vector<float> w((size_t)cout * cin * taps), scale(cout), shift(cout);
    int co = 0;
    for (const Member& m : ms) {
      auto wt = net->raw.find(m.name + ".conv.weight");
      if (wt == net->raw.end() || wt->second.size() != (size_t)m.out.c * cin * taps) {
        rc = fail(ctx, TISE_ERR_INVALID, "state_dict: missing or mis-sized %s.conv.weight (want %dx%dx%dx%d)",
                  m.name.c_str(), m.out.c, cin, R, S);
        return nullptr;
      }
      std::copy(wt->second.begin(), wt->second.end(), w.begin() + (size_t)co * cin * taps);
      std::vector<float> sc, sh;
      if ((rc = fold_bn(ctx, net, m.name, m.out.c, &sc, &sh))) return nullptr;
      if (m.raw) {  // conv output stays un-normalised; keep the BN for the pooling step
        PoolBN pb;
        if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&pb.d_scale, sc.size() * 4))) return nullptr;
        if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&pb.d_shift, sh.size() * 4))) return nullptr;
        cudaMemcpy(pb.d_scale, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(pb.d_shift, sh.data(), sh.size() * 4, cudaMemcpyHostToDevice);
        net->pool_bn[m.name] = pb;
        std::fill(sc.begin(), sc.end(), 1.0f);
        std::fill(sh.begin(), sh.end(), 0.0f);
      }
      std::copy(sc.begin(), sc.end(), scale.begin() + co);
      std::copy(sh.begin(), sh.end(), shift.begin() + co);
      co += m.out.c;
    }
    ConvLayer L;
    L.cin = cin;
    L.cout = cout;
    L.R = R;
    L.S = S;
    L.stride = stride;
    L.pad_h = ph;
    L.pad_w = pw;
    if (fold_s) {  // w[co][ci][r][s] -> w'[co][s * cin + ci][r][0]
      std::vector<float> w2(w.size());
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
          for (int r = 0; r < R; ++r)
            for (int sx = 0; sx < S; ++sx)
              w2[((size_t)co * (S * cin) + (size_t)sx * cin + ci) * R + r] = w[(((size_t)co * cin + ci) * R + r) * S + sx];
      w.swap(w2);
      L.cin = S * cin;
      L.S = 1;
      L.fold = S;
    }
    if ((rc = pack_conv_layer(ctx, &L, w.data(), scale.data(), shift.data(), &net->device_allocs, force_kc))) return nullptr;
    return &(net->convs[key] = L);
  }

  void conv_multi(const std::string& key, View in, const std::vector<Member>& ms, int R = 1, int S = 1, int stride = 1,
                  int ph = 0, int pw = 0) {
    if (rc) return;
    const Buf& bi = bufs[in.buf];
    const ConvLayer* L = layer(key, ms, in.c, R, S, stride, ph, pw);
    if (!L) return;
    const int ho = conv_out_dim(bi.h, R, stride, ph), wo = conv_out_dim(bi.w, S, stride, pw);
    for (const Member& m : ms)
      if (bufs[m.out.buf].h != ho || bufs[m.out.buf].w != wo || bufs[m.out.buf].f32 != bufs[ms[0].out.buf].f32) {
        rc = fail(ctx, TISE_ERR_INVALID, "internal: %s output %dx%d does not match its buffer", m.name.c_str(), ho, wo);
        return;
      }
    if (!plan) return;
    Step s;
    s.kind = STEP_CONV;
    s.layer = L;
    s.conv.in = (const __half*)ptr(in);
    s.conv.n = n;
    s.conv.h = bi.h;
    s.conv.w = bi.w;
    s.conv.in_stride = bi.c;
    s.conv.n_segs = (int)ms.size();
    for (size_t i = 0; i < ms.size(); ++i) {
      s.conv.segs[i].out = ptr(ms[i].out);
      s.conv.segs[i].out_stride = bufs[ms[i].out.buf].c;
      s.conv.segs[i].c = ms[i].out.c;
      s.conv.segs[i].relu = ms[i].raw ? 0 : 1;
    }
    s.conv.out_f32 = bufs[ms[0].out.buf].f32;
    s.conv.ho = ho;
    s.conv.wo = wo;
    plan->steps.push_back(s);
  }
  void conv(const std::string& name, View in, View out, int R = 1, int S = 1, int stride = 1, int ph = 0, int pw = 0) {
    conv_multi(name, in, {Member{name, out, false}}, R, S, stride, ph, pw);
  }
  // stride-1, unpadded R x S convolution evaluated as an R x 1 convolution over windows of S pixels (see layer())
  void conv_folded(const std::string& name, View in, View out, int R, int S) {
    if (rc) return;
    const Buf& bi = bufs[in.buf];
    const ConvLayer* L = layer(name, {Member{name, out, false}}, in.c, R, S, 1, 0, 0, 0, true);
    if (!L) return;
    const int ho = bi.h - R + 1, wo = bi.w - S + 1;
    if (bufs[out.buf].h != ho || bufs[out.buf].w != wo || in.choff != 0 || in.c != bi.c) {
      rc = fail(ctx, TISE_ERR_INVALID, "internal: folded conv %s does not match its buffers", name.c_str());
      return;
    }
    if (!plan) return;
    Step s;
    s.kind = STEP_CONV;
    s.layer = L;
    s.conv.in = (const __half*)ptr(in);
    s.conv.n = n;
    s.conv.h = bi.h;
    s.conv.w = wo;  // logical width: one "pixel" per output column
    s.conv.in_row_pitch = bi.w;
    s.conv.in_stride = bi.c;
    s.conv.n_segs = 1;
    s.conv.segs[0].out = ptr(out);
    s.conv.segs[0].out_stride = bufs[out.buf].c;
    s.conv.segs[0].c = out.c;
    s.conv.segs[0].relu = 1;
    s.conv.out_f32 = bufs[out.buf].f32;
    s.conv.ho = ho;
    s.conv.wo = wo;
    plan->steps.push_back(s);
  }
  // flat convolution `name` from buffer `in` to buffer `out` (same physical grid); valid outputs ho x wo
  void flat(const std::string& name, int in, int out, int ho, int wo, int out_shift) {
    if (rc || !plan) return;
    const Buf& bi = bufs[in];
    Step z;  // the cells in front of the first shifted output are never written: keep them zero
    z.kind = STEP_ZERO;
    z.out = ptr(all(out));
    z.zero_bytes = (size_t)out_shift * bufs[out].c * 2;
    if (z.zero_bytes) plan->steps.push_back(z);
    Step s;
    s.kind = STEP_FLATCONV;
    s.layer = &net->convs[name];
    s.flat.in = (const __half*)ptr(all(in));
    s.flat.out = (__half*)ptr(all(out));
    s.flat.n = n;
    s.flat.hp = bi.h;
    s.flat.wp = bi.w;
    s.flat.ho = ho;
    s.flat.wo = wo;
    s.flat.out_shift = out_shift;
    s.flat.relu = 1;
    plan->steps.push_back(s);
  }
  void pool(StepKind kind, View in, View out, int logical_h = 0, int logical_w = 0) {
    if (rc || !plan) return;
    const Buf& bi = bufs[in.buf];
    const Buf& bo = bufs[out.buf];
    Step s;
    s.kind = kind;
    s.in = ptr(in);
    s.out = ptr(out);
    s.n = n;
    s.h = logical_h ? logical_h : bi.h;
    s.w = logical_w ? logical_w : bi.w;
    s.in_rows = bi.h;
    s.in_pitch = bi.w;
    s.c = in.c;
    s.in_stride = bi.c;
    s.out_stride = bo.c;
    plan->steps.push_back(s);
  }
  // avg_pool2d(3,1,1) + BN + ReLU of the raw branch_pool convolution `name` (buffer `in`, all channels) -> out
  void pool_bn(const std::string& name, int in, View out) {
    if (rc || !plan) return;
    const Buf& bi = bufs[in];
    const Buf& bo = bufs[out.buf];
    const PoolBN& pb = net->pool_bn[name];
    Step s;
    s.kind = STEP_AVGPOOL_BN;
    s.in = ptr(all(in));
    s.out = ptr(out);
    s.n = n;
    s.h = bi.h;
    s.w = bi.w;
    s.c = bi.c;
    s.in_stride = bi.c;
    s.out_stride = bo.c;
    s.out_f32 = bo.f32;
    s.pool_scale = pb.d_scale;
    s.pool_shift = pb.d_shift;
    plan->steps.push_back(s);
  }
  void gap(View in) {
    if (rc || !plan) return;
    const Buf& bi = bufs[in.buf];
    Step s;
    s.kind = STEP_GAP;
    s.in = ptr(in);
    s.n = n;
    s.h = bi.h;
    s.w = bi.w;
    s.c = in.c;
    s.in_stride = bi.c;
    s.in_f32 = bi.f32;
    plan->steps.push_back(s);
  }

  // In every block the 1x1 convolutions that read the block input run as ONE GEMM (one A-operand pass, wider N),
  // and branch_pool's avg_pool -> conv1x1 -> BN -> ReLU is evaluated as conv1x1 (raw) -> avg_pool -> BN -> ReLU,
  // which is the same function (both linear, zero padding, no bias) on pf instead of cin channels.

  // torchvision InceptionA (models/inception.py): cat[branch1x1 64, branch5x5 64, branch3x3dbl 96, branch_pool pf]
  int inceptionA(const std::string& p, int in, int pf) {
    const int hw = bufs[in].h;
    const int out = buf(hw, hw, 224 + pf);
    const int t5 = buf(hw, hw, 48), d1 = buf(hw, hw, 64), d2 = buf(hw, hw, 96), tp = buf(hw, hw, pf);
    conv_multi(p + ".fused1x1", all(in),
               {Member{p + ".branch1x1", View{out, 0, 64}, false}, Member{p + ".branch5x5_1", all(t5), false},
                Member{p + ".branch3x3dbl_1", all(d1), false}, Member{p + ".branch_pool", all(tp), true}});
    conv(p + ".branch5x5_2", all(t5), View{out, 64, 64}, 5, 5, 1, 2, 2);
    conv(p + ".branch3x3dbl_2", all(d1), all(d2), 3, 3, 1, 1, 1);
    conv(p + ".branch3x3dbl_3", all(d2), View{out, 128, 96}, 3, 3, 1, 1, 1);
    pool_bn(p + ".branch_pool", tp, View{out, 224, pf});
    return out;
  }
  // InceptionB: cat[branch3x3 384 (s2), branch3x3dbl 96 (s2), maxpool cin]
  int inceptionB(const std::string& p, int in) {
    const int hw = bufs[in].h, cin = bufs[in].c;
    const int ho = (hw - 3) / 2 + 1;
    const int out = buf(ho, ho, 384 + 96 + cin);
    conv(p + ".branch3x3", all(in), View{out, 0, 384}, 3, 3, 2, 0, 0);
    const int d1 = buf(hw, hw, 64), d2 = buf(hw, hw, 96);
    conv(p + ".branch3x3dbl_1", all(in), all(d1));
    conv(p + ".branch3x3dbl_2", all(d1), all(d2), 3, 3, 1, 1, 1);
    conv(p + ".branch3x3dbl_3", all(d2), View{out, 384, 96}, 3, 3, 2, 0, 0);
    pool(STEP_MAXPOOL, all(in), View{out, 480, cin});
    return out;
  }
  // InceptionC: cat[branch1x1 192, branch7x7 192, branch7x7dbl 192, branch_pool 192]
  int inceptionC(const std::string& p, int in, int c7) {
    const int hw = bufs[in].h;
    const int out = buf(hw, hw, 768);
    const int a1 = buf(hw, hw, c7), a2 = buf(hw, hw, c7);
    const int b1 = buf(hw, hw, c7), b2 = buf(hw, hw, c7), b3 = buf(hw, hw, c7), b4 = buf(hw, hw, c7);
    const int tp = buf(hw, hw, 192);
    conv_multi(p + ".fused1x1", all(in),
               {Member{p + ".branch1x1", View{out, 0, 192}, false}, Member{p + ".branch7x7_1", all(a1), false},
                Member{p + ".branch7x7dbl_1", all(b1), false}, Member{p + ".branch_pool", all(tp), true}});
    conv(p + ".branch7x7_2", all(a1), all(a2), 1, 7, 1, 0, 3);
    conv(p + ".branch7x7_3", all(a2), View{out, 192, 192}, 7, 1, 1, 3, 0);
    conv(p + ".branch7x7dbl_2", all(b1), all(b2), 7, 1, 1, 3, 0);
    conv(p + ".branch7x7dbl_3", all(b2), all(b3), 1, 7, 1, 0, 3);
    conv(p + ".branch7x7dbl_4", all(b3), all(b4), 7, 1, 1, 3, 0);
    conv(p + ".branch7x7dbl_5", all(b4), View{out, 384, 192}, 1, 7, 1, 0, 3);
    pool_bn(p + ".branch_pool", tp, View{out, 576, 192});
    return out;
  }
  // InceptionD: cat[branch3x3 320 (s2), branch7x7x3 192 (s2), maxpool cin]
  int inceptionD(const std::string& p, int in) {
    const int hw = bufs[in].h, cin = bufs[in].c;
    const int ho = (hw - 3) / 2 + 1;
    const int out = buf(ho, ho, 320 + 192 + cin);
    const int a1 = buf(hw, hw, 192), b1 = buf(hw, hw, 192), b2 = buf(hw, hw, 192), b3 = buf(hw, hw, 192);
    conv_multi(p + ".fused1x1", all(in),
               {Member{p + ".branch3x3_1", all(a1), false}, Member{p + ".branch7x7x3_1", all(b1), false}});
    conv(p + ".branch3x3_2", all(a1), View{out, 0, 320}, 3, 3, 2, 0, 0);
    conv(p + ".branch7x7x3_2", all(b1), all(b2), 1, 7, 1, 0, 3);
    conv(p + ".branch7x7x3_3", all(b2), all(b3), 7, 1, 1, 3, 0);
    conv(p + ".branch7x7x3_4", all(b3), View{out, 320, 192}, 3, 3, 2, 0, 0);
    pool(STEP_MAXPOOL, all(in), View{out, 512, cin});
    return out;
  }
  // InceptionE: cat[branch1x1 320, branch3x3 (2a|2b) 768, branch3x3dbl (3a|3b) 768, branch_pool 192]
  int inceptionE(const std::string& p, int in, int out_f32) {
    const int hw = bufs[in].h;
    const int out = buf(hw, hw, 2048, out_f32);
    const int a1 = buf(hw, hw, 384), b1 = buf(hw, hw, 448), b2 = buf(hw, hw, 384), tp = buf(hw, hw, 192);
    if (out_f32) {  // the fp32 destination of branch1x1 cannot share a launch with the fp16 temporaries
      conv(p + ".branch1x1", all(in), View{out, 0, 320});
      conv_multi(p + ".fused1x1", all(in),
                 {Member{p + ".branch3x3_1", all(a1), false}, Member{p + ".branch3x3dbl_1", all(b1), false},
                  Member{p + ".branch_pool", all(tp), true}});
    } else {
      conv_multi(p + ".fused1x1", all(in),
                 {Member{p + ".branch1x1", View{out, 0, 320}, false}, Member{p + ".branch3x3_1", all(a1), false},
                  Member{p + ".branch3x3dbl_1", all(b1), false}, Member{p + ".branch_pool", all(tp), true}});
    }
    conv(p + ".branch3x3_2a", all(a1), View{out, 320, 384}, 1, 3, 1, 0, 1);
    conv(p + ".branch3x3_2b", all(a1), View{out, 704, 384}, 3, 1, 1, 1, 0);
    conv(p + ".branch3x3dbl_2", all(b1), all(b2), 3, 3, 1, 1, 1);
    conv(p + ".branch3x3dbl_3a", all(b2), View{out, 1088, 384}, 1, 3, 1, 0, 1);
    conv(p + ".branch3x3dbl_3b", all(b2), View{out, 1472, 384}, 3, 1, 1, 1, 0);
    pool_bn(p + ".branch_pool", tp, View{out, 1856, 192});
    return out;
  }

  void walk(int tap) {
    // block 0
    const int c1 = buf(149, 149, 32);
    if (plan) {
      Step s;
      s.kind = STEP_STEM;
      s.out = ptr(all(c1));
      s.n = n;
      plan->steps.push_back(s);
    }
    // Conv2d_2a / 2b have 32 input channels: packed with 32-channel K chunks and, on the tcgen05 path, evaluated by
    // the flat kernel on the 149 x 149 grid (2a's output is written one row + one pixel further, which makes it the
    // zero-padded input of the pad-1 convolution 2b; 2b's output keeps the 149-pixel pitch for the max pool)
    layer("Conv2d_2a_3x3", {Member{"Conv2d_2a_3x3", View{0, 0, 32}, false}}, 32, 3, 3, 1, 0, 0, 32);
    layer("Conv2d_2b_3x3", {Member{"Conv2d_2b_3x3", View{0, 0, 64}, false}}, 32, 3, 3, 1, 1, 1, 32);
    const int p1 = buf(73, 73, 64);
    if (flat_stem) {
      const int c2 = buf(149, 149, 32), c3 = buf(149, 149, 64);
      flat("Conv2d_2a_3x3", c1, c2, 147, 147, 149 + 1);
      flat("Conv2d_2b_3x3", c2, c3, 147, 147, 0);
      pool(STEP_MAXPOOL, all(c3), all(p1), 147, 147);
    } else {
      const int c2 = buf(147, 147, 32), c3 = buf(147, 147, 64);
      conv("Conv2d_2a_3x3", all(c1), all(c2), 3, 3, 1, 0, 0);
      conv("Conv2d_2b_3x3", all(c2), all(c3), 3, 3, 1, 1, 1);
      pool(STEP_MAXPOOL, all(c3), all(p1));
    }
    if (tap == TISE_TAP_64) return gap(all(p1));
    // block 1
    const int c4 = buf(73, 73, 80), c5 = buf(71, 71, 192), p2 = buf(35, 35, 192);
    conv("Conv2d_3b_1x1", all(p1), all(c4));
    static int fold_env = -1;
    if (fold_env < 0) fold_env = getenv("TISE_FOLD_4A") ? atoi(getenv("TISE_FOLD_4A")) : 1;
    if (fold_env) conv_folded("Conv2d_4a_3x3", all(c4), all(c5), 3, 3);
    else conv("Conv2d_4a_3x3", all(c4), all(c5), 3, 3, 1, 0, 0);
    pool(STEP_MAXPOOL, all(c5), all(p2));
    if (tap == TISE_TAP_192) return gap(all(p2));
    // block 2
    int x = inceptionA("Mixed_5b", p2, 32);
    x = inceptionA("Mixed_5c", x, 64);
    x = inceptionA("Mixed_5d", x, 64);
    x = inceptionB("Mixed_6a", x);
    x = inceptionC("Mixed_6b", x, 128);
    x = inceptionC("Mixed_6c", x, 160);
    x = inceptionC("Mixed_6d", x, 160);
    x = inceptionC("Mixed_6e", x, 192);
    if (tap == TISE_TAP_768) return gap(all(x));
    // block 3
    x = inceptionD("Mixed_7a", x);
    x = inceptionE("Mixed_7b", x, 0);
    x = inceptionE("Mixed_7c", x, 1);  // fp32 output feeding the global average pool (pool3)
    gap(all(x));
  }
};

// ---------------------------------------------------------------------------------------------
int net_load(tise_ctx* ctx, const tise_weight_desc* tensors, int n_tensors, int num_classes) {
  net_free(ctx);
  Net* net = new Net();
  ctx->net = net;
  for (int i = 0; i < n_tensors; ++i) {
    if (!tensors[i].name || !tensors[i].data || tensors[i].numel < 0)
      return fail(ctx, TISE_ERR_INVALID, "tise_load_inception: tensor %d is null", i);
    net->raw[tensors[i].name].assign(tensors[i].data, tensors[i].data + tensors[i].numel);
  }
  // stem: Conv2d_1a_3x3 kept in fp32, layout [k = (r*3+s)*3+c][co]
  {
    auto wt = net->raw.find("Conv2d_1a_3x3.conv.weight");
    if (wt == net->raw.end() || wt->second.size() != 32 * 27)
      return fail(ctx, TISE_ERR_INVALID, "state_dict: missing or mis-sized Conv2d_1a_3x3.conv.weight");
    std::vector<float> scale, shift, w(27 * 32);
    int rc = fold_bn(ctx, net, "Conv2d_1a_3x3", 32, &scale, &shift);
    if (rc) return rc;
    for (int co = 0; co < 32; ++co)
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s) w[((r * 3 + s) * 3 + c) * 32 + co] = wt->second[((co * 3 + c) * 3 + r) * 3 + s];
    if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&net->d_stem_w, w.size() * 4))) return rc;
    if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&net->d_stem_scale, 32 * 4))) return rc;
    if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&net->d_stem_shift, 32 * 4))) return rc;
    TISE_CUDA(ctx, cudaMemcpy(net->d_stem_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    TISE_CUDA(ctx, cudaMemcpy(net->d_stem_scale, scale.data(), 32 * 4, cudaMemcpyHostToDevice));
    TISE_CUDA(ctx, cudaMemcpy(net->d_stem_shift, shift.data(), 32 * 4, cudaMemcpyHostToDevice));
    std::vector<__half> w16(64 * 64);
    std::vector<float> scale16(scale);
    pack_stem_umma(w.data(), scale16.data(), w16.data());
    if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&net->d_stem_w16, w16.size() * 2))) return rc;
    if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&net->d_stem_scale16, 32 * 4))) return rc;
    TISE_CUDA(ctx, cudaMemcpy(net->d_stem_w16, w16.data(), w16.size() * 2, cudaMemcpyHostToDevice));
    TISE_CUDA(ctx, cudaMemcpy(net->d_stem_scale16, scale16.data(), 32 * 4, cudaMemcpyHostToDevice));
  }
  // declare-only walks: create every layer, size the workspace per tap
  for (int tap = 0; tap < 4; ++tap) {
    Builder b{ctx, net, nullptr, 0, nullptr};
    b.walk(tap);
    if (b.rc) return b.rc;
    net->bytes_per_image[tap] = b.per_image;
  }
  // classifier head
  net->num_classes = num_classes;
  if (num_classes > 0) {
    auto fw = net->raw.find("fc.weight");
    auto fb = net->raw.find("fc.bias");
    if (fw == net->raw.end() || fw->second.size() != (size_t)num_classes * 2048)
      return fail(ctx, TISE_ERR_INVALID, "state_dict: missing or mis-sized fc.weight (want %dx2048)", num_classes);
    int rc;
    if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&net->d_fc_w, fw->second.size() * 4))) return rc;
    if ((rc = dev_alloc(ctx, &net->device_allocs, (void**)&net->d_fc_b, (size_t)num_classes * 4))) return rc;
    TISE_CUDA(ctx, cudaMemcpy(net->d_fc_w, fw->second.data(), fw->second.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> bias(num_classes, 0.f);
    if (fb != net->raw.end() && (int)fb->second.size() == num_classes) bias = fb->second;
    TISE_CUDA(ctx, cudaMemcpy(net->d_fc_b, bias.data(), (size_t)num_classes * 4, cudaMemcpyHostToDevice));
  }
  net->raw.clear();
  TISE_CUDA(ctx, cudaDeviceSynchronize());
  return TISE_OK;
}

void net_free(tise_ctx* ctx) {
  if (!ctx->net) return;
  for (void* p : ctx->net->device_allocs) cudaFree(p);
  delete ctx->net;
  ctx->net = nullptr;
}

static int get_plan(tise_ctx* ctx, void* ws, int n, int tap, Plan** out) {
  Net* net = ctx->net;
  static int flat_env = -1;
  if (flat_env < 0) flat_env = getenv("TISE_FLAT_STEM") ? atoi(getenv("TISE_FLAT_STEM")) : 1;
  const bool flat_stem = ctx->conv_impl == TISE_CONV_UMMA && ctx->passes == 2 && flat_env != 0;
  for (auto& p : net->plans)
    if (p->ws == ws && p->n == n && p->tap == tap && p->impl == (flat_stem ? 1 : 0)) {
      *out = p.get();
      return TISE_OK;
    }
  std::unique_ptr<Plan> plan(new Plan());
  plan->ws = ws;
  plan->n = n;
  plan->tap = tap;
  plan->impl = flat_stem ? 1 : 0;
  Builder b{ctx, net, plan.get(), n, (char*)ws};
  b.flat_stem = flat_stem;
  b.walk(tap);
  if (b.rc) return b.rc;
  if (net->plans.size() >= 16) net->plans.erase(net->plans.begin());
  net->plans.push_back(std::move(plan));
  *out = net->plans.back().get();
  return TISE_OK;
}

// ---------------------------------------------------------------------------------------------
// Single-pass calibration ("mean-field W_lo").  A layer run with W_hi only is off by  sum_k W_lo[co][k] a_k, whose
// image-independent part  sum_k W_lo[co][k] mean(a_k)  is what moves the feature MEAN (and with it IS and any FID
// against statistics computed elsewhere); the rest is zero-mean noise of the size of the fp16 activation rounding.
// The channel means of every convolution input are measured once, on a fixed set of synthetic images generated here
// (they are a property of the network's BatchNorm shifts far more than of the images: calibrating on other image
// distributions, even white noise, gives the same result to ~1e-5 -- tools/bias_correction_probe.py), and folded into
// the shift of the single-pass form of the layer.  Nothing here depends on the caller's data, so results stay a pure
// function of (weights, images).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t calib_hash(uint32_t h) {
  h ^= h >> 15;
  h *= 0x85EBCA77u;
  h ^= h >> 13;
  h *= 0xC2B2AE3Du;
  h ^= h >> 16;
  return h;
}
// smooth random colour fields (5 x 5 control grid, bilinear) with per-image gain and pixel noise: uint8 [n][299][299][3]
__global__ void calib_images_kernel(uint8_t* out, int n) {
  const size_t total = (size_t)n * kOutSize * kOutSize * 3;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 3);
    size_t p = i / 3;
    const int x = (int)(p % kOutSize);
    p /= kOutSize;
    const int y = (int)(p % kOutSize), img = (int)(p / kOutSize);
    const int fy = y * 4 * 256 / kOutSize, fx = x * 4 * 256 / kOutSize;  // 8.8 fixed point, 4 cells across
    const int iy = min(fy >> 8, 3), ix = min(fx >> 8, 3), ay = fy - (iy << 8), ax = fx - (ix << 8);
    auto g = [&](int gy, int gx) { return (int)(calib_hash((uint32_t)(((img * 5 + gy) * 5 + gx) * 3 + c) + 0x9E3779B1u) & 255u); };
    const int top = g(iy, ix) * (256 - ax) + g(iy, ix + 1) * ax, bot = g(iy + 1, ix) * (256 - ax) + g(iy + 1, ix + 1) * ax;
    const int v = (top * (256 - ay) + bot * ay) >> 16;
    const int gain = 128 + (int)(calib_hash((uint32_t)img * 2654435761u + 77u) & 255u);
    const int noise = (int)(calib_hash((uint32_t)i + 0x7F4A7C15u) & 63u) - 32;
    out[i] = (uint8_t)max(0, min(255, ((v * gain) >> 8) + noise));
  }
}
// sum[c] += sum over pixels of in[pixel * stride + c]   (32 channels x 8 pixel lanes per CTA)
__global__ void __launch_bounds__(256) channel_sum_kernel(const __half* __restrict__ in, long long pixels, int stride, int c_total,
                                                          double* __restrict__ sum) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if (c >= c_total) return;
  double acc = 0.0;
  for (long long p = (long long)blockIdx.y * 8 + (threadIdx.x >> 5); p < pixels; p += (long long)gridDim.y * 8)
    acc += (double)__half2float(in[p * stride + c]);
  atomicAdd(sum + c, acc);
}
// shift1[co] = shift[co] + scale[co] * sum_k W_lo[co][k] * mean[channel of k] * valid[tap of k]   (one warp per output channel)
// valid[tap]: fraction of the layer's output positions for which that filter tap reads a real pixel rather than zero
// padding (17 x 17 maps with 1x7 filters: 14/17 for the outermost taps, 8 x 8 maps with 3x3: 7/8 per padded direction).
// Without it the constant over-corrects every border position, which left ~10 % of the uncorrected mean shift in place.
struct TapValid {
  float v[64];
};
__global__ void __launch_bounds__(256) lo_correction_kernel(const __half* __restrict__ w, int ktot, int cin_pad, int cin, int cin_phys,
                                                            int nt, int cout, const double* __restrict__ sum, double inv_pixels,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            float* __restrict__ shift1, const TapValid valid) {
  const int co = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (co >= cout) return;
  const __half* lo = w + (((size_t)(co / nt) * 2 + 1) * nt + co % nt) * ktot;
  double acc = 0.0;
  for (int k = lane; k < ktot; k += 32) {
    const int cc = k % cin_pad, tap = k / cin_pad;
    if (cc < cin) acc += (double)__half2float(lo[k]) * (sum[cc % cin_phys] * inv_pixels) * (double)valid.v[tap < 64 ? tap : 63];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) shift1[co] = shift[co] + scale[co] * (float)acc;
}

static int get_plan(tise_ctx* ctx, void* ws, int n, int tap, Plan** out);

int net_calibrate_lo(tise_ctx* ctx, int n, void* d_ws, size_t ws_bytes, cudaStream_t st) {
  Net* net = ctx->net;
  if (!net) return fail(ctx, TISE_ERR_STATE, "tise_calibrate_single_pass: weights not loaded");
  if (!ctx->lut_set) return fail(ctx, TISE_ERR_STATE, "tise_calibrate_single_pass: input LUT not set");
  if (n < 1 || n > 4096) return fail(ctx, TISE_ERR_INVALID, "tise_calibrate_single_pass: n = %d", n);
  uint8_t* d_img = nullptr;
  float* d_feat = nullptr;
  double* d_sum = nullptr;
  TISE_CUDA(ctx, cudaMalloc(&d_img, (size_t)n * kOutSize * kOutSize * 3));
  TISE_CUDA(ctx, cudaMalloc(&d_feat, (size_t)n * 2048 * sizeof(float)));
  TISE_CUDA(ctx, cudaMalloc(&d_sum, 4096 * sizeof(double)));
  calib_images_kernel<<<ctx->num_sms * 8, 256, 0, st>>>(d_img, n);
  // strict forward: every activation tensor of the graph ends up in the workspace (no buffer is reused)
  const int keep_passes = ctx->passes, keep_hw = ctx->pass1_max_hw, keep_impl = ctx->conv_impl;
  const bool keep_prof = ctx->profiling;
  ctx->passes = 2;
  ctx->pass1_max_hw = 0;
  ctx->conv_impl = TISE_CONV_UMMA;
  ctx->profiling = false;
  int rc = net_forward(ctx, d_img, nullptr, n, TISE_TAP_2048, d_feat, d_ws, ws_bytes, st);
  Plan* plan = nullptr;
  if (!rc) rc = get_plan(ctx, d_ws, n, TISE_TAP_2048, &plan);
  ctx->passes = keep_passes;
  ctx->pass1_max_hw = keep_hw;
  ctx->conv_impl = keep_impl;
  ctx->profiling = keep_prof;
  if (!rc) {
    for (Step& s : plan->steps) {
      if (s.kind != STEP_CONV && s.kind != STEP_FLATCONV) continue;
      const ConvLayer& L = *s.layer;
      const int cin_phys = L.cin / L.fold;
      const bool flat = s.kind == STEP_FLATCONV;  // flat layers read the whole padded grid (its zero cells dilute the
      const __half* in = flat ? s.flat.in : s.conv.in;  // mean by (147 / 149)^2: second order, like the border taps)
      const int stride = flat ? L.cin : s.conv.in_stride;
      const long long pixels = flat ? (long long)s.flat.n * s.flat.hp * s.flat.wp
                                    : (long long)s.conv.n * s.conv.h * (s.conv.in_row_pitch ? s.conv.in_row_pitch : s.conv.w);
      cudaMemsetAsync(d_sum, 0, 4096 * sizeof(double), st);
      int slices = (int)((pixels + 8 * 64 - 1) / (8 * 64));
      if (slices > 1024) slices = 1024;
      if (slices < 1) slices = 1;
      channel_sum_kernel<<<dim3((cin_phys + 31) / 32, slices), 256, 0, st>>>(in, pixels, stride, cin_phys, d_sum);
      TapValid tv;
      for (int t = 0; t < 64; ++t) tv.v[t] = 1.0f;
      static int tv_env = -1;
      if (tv_env < 0) tv_env = getenv("TISE_LO_TAPVALID") ? atoi(getenv("TISE_LO_TAPVALID")) : 1;
      if (!flat && tv_env && L.R * L.S <= 64) {  // (flat layers: their padding is part of the grid the mean was taken over)
        auto frac = [](int n_in, int n_out, int str, int pad, int tap) {  // output positions whose tap lands inside [0, n_in)
          int ok = 0;
          for (int o = 0; o < n_out; ++o) {
            const int i = o * str - pad + tap;
            ok += i >= 0 && i < n_in;
          }
          return n_out > 0 ? (float)ok / (float)n_out : 1.0f;
        };
        for (int r = 0; r < L.R; ++r)
          for (int sx = 0; sx < L.S; ++sx)
            tv.v[r * L.S + sx] = frac(s.conv.h, s.conv.ho, L.stride, L.pad_h, r) *
                                 (L.fold > 1 ? 1.0f : frac(s.conv.w, s.conv.wo, L.stride, L.pad_w, sx));
      }
      lo_correction_kernel<<<(L.cout + 7) / 8, 256, 0, st>>>(L.d_w, L.ktot, L.cin_pad, L.cin, cin_phys, L.nt, L.cout, d_sum,
                                                            1.0 / (double)pixels, L.d_scale, L.d_shift, L.d_shift1, tv);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(ctx, TISE_ERR_CUDA, "tise_calibrate_single_pass: %s", cudaGetErrorString(e));
  }
  cudaFree(d_img);
  cudaFree(d_feat);
  cudaFree(d_sum);
  return rc;
}

int net_forward(tise_ctx* ctx, const uint8_t* d_u8, const float* d_f32, int n, int tap, float* d_feat, void* d_ws,
                size_t ws_bytes, cudaStream_t st) {
  Net* net = ctx->net;
  if (!net) return fail(ctx, TISE_ERR_STATE, "tise_inception_features: weights not loaded");
  if (tap < 0 || tap > 3) return fail(ctx, TISE_ERR_INVALID, "tise_inception_features: bad tap %d", tap);
  if (n == 0) return TISE_OK;
  if (n < 0 || (!d_u8 == !d_f32) || !d_feat || !d_ws)
    return fail(ctx, TISE_ERR_INVALID, "tise_inception_features: need exactly one input, an output and a workspace");
  if (d_u8 && !ctx->lut_set) return fail(ctx, TISE_ERR_STATE, "tise_inception_features: input LUT not set");
  const size_t need = net->bytes_per_image[tap] * (size_t)n;
  if (ws_bytes < need)
    return fail(ctx, TISE_ERR_WORKSPACE, "tise_inception_features: workspace %zu < %zu bytes", ws_bytes, need);
  if ((uintptr_t)d_ws & 255) return fail(ctx, TISE_ERR_INVALID, "tise_inception_features: workspace must be 256-byte aligned");
  Plan* plan = nullptr;
  int rc = get_plan(ctx, d_ws, n, tap, &plan);
  if (rc) return rc;
  static int stem_env = -1;
  if (stem_env < 0) stem_env = getenv("TISE_STEM_UMMA") ? atoi(getenv("TISE_STEM_UMMA")) : 1;
  const bool stem_umma = ctx->conv_impl == TISE_CONV_UMMA && ctx->passes == 2 && stem_env != 0;
  bool progress_pending = ctx->progress_event != nullptr;
  for (Step& s : plan->steps) {
    if (progress_pending && s.kind == STEP_CONV && s.conv.h <= ctx->progress_max_hw) {  // tise_set_progress_event
      TISE_CUDA(ctx, cudaEventRecord(ctx->progress_event, st));
      progress_pending = false;
    }
    tise_ctx::ProfRec rec{s.kind == STEP_FLATCONV ? (int)STEP_CONV : (int)s.kind, nullptr, nullptr, 0.0, {0}};
    if (ctx->profiling) {
      if (s.kind == STEP_FLATCONV) {
        snprintf(rec.tag, sizeof(rec.tag), "conv n%d %dx%d cin%d cout%d k%dx%d s1 flat", s.flat.n, s.flat.hp, s.flat.wp,
                 s.layer->cin, s.layer->cout, s.layer->R, s.layer->S);
        rec.flops = 2.0 * (double)s.flat.n * s.flat.ho * s.flat.wo * s.layer->cout * s.layer->R * s.layer->S * s.layer->cin;
      } else if (s.kind == STEP_CONV)
        snprintf(rec.tag, sizeof(rec.tag), "conv n%d %dx%d cin%d cout%d k%dx%d s%d", s.conv.n, s.conv.h, s.conv.w,
                 s.layer->cin, s.layer->cout, s.layer->R, s.layer->S, s.layer->stride);
      else
        snprintf(rec.tag, sizeof(rec.tag), "kind%d n%d %dx%d c%d", (int)s.kind, s.n, s.h, s.w, s.c);
      TISE_CUDA(ctx, cudaEventCreate(&rec.a));
      TISE_CUDA(ctx, cudaEventCreate(&rec.b));
      if (s.kind == STEP_CONV)  // algorithmic FLOPs: 2 * M * Cout * (R * S * Cin), padding channels excluded
        rec.flops = 2.0 * (double)s.conv.n * s.conv.ho * s.conv.wo * s.layer->cout * s.layer->R * s.layer->S * s.layer->cin;
      if (s.kind == STEP_STEM) rec.flops = 2.0 * (double)n * 149 * 149 * 32 * 27;
      TISE_CUDA(ctx, cudaEventRecord(rec.a, st));
    }
    switch (s.kind) {
      case STEP_STEM:
        if (d_u8 && stem_umma)  // tensor-core stem (uint8 input, tcgen05 engine, two passes); else the fp32 CUDA-core kernel
          rc = launch_stem_umma(ctx, d_u8, n, net->d_stem_w16, net->d_stem_scale16, net->d_stem_shift, (__half*)s.out, st);
        else
          rc = launch_stem(ctx, d_u8, d_f32, n, net->d_stem_w, net->d_stem_scale, net->d_stem_shift, (__half*)s.out, st);
        break;
      case STEP_CONV:
        if (ctx->conv_impl == TISE_CONV_UMMA) {
          if (!s.umma.valid && (rc = build_umma_plan(ctx, *s.layer, s.conv, &s.umma))) break;
          rc = launch_conv_umma(ctx, *s.layer, s.conv, s.umma,
                                (ctx->passes == 2 && s.conv.h <= ctx->pass1_max_hw) ? 1 : ctx->passes, st);
        } else {
          rc = launch_conv_simt(ctx, *s.layer, s.conv, st);
        }
        break;
      case STEP_MAXPOOL:
        rc = launch_maxpool3s2(ctx, (const __half*)s.in, s.n, s.h, s.w, s.c, s.in_stride, (__half*)s.out, s.out_stride,
                               st, s.in_rows, s.in_pitch);
        break;
      case STEP_FLATCONV:
        if (!s.flatplan.valid && (rc = build_flat_plan(ctx, *s.layer, s.flat, &s.flatplan))) break;
        rc = launch_conv_flat(ctx, *s.layer, s.flat, s.flatplan, (s.flat.hp <= ctx->pass1_max_hw) ? 1 : 2, st);
        break;
      case STEP_ZERO: {
        cudaError_t e = cudaMemsetAsync(s.out, 0, s.zero_bytes, st);
        if (e != cudaSuccess) rc = fail(ctx, TISE_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        break;
      }
      case STEP_AVGPOOL:
        rc = launch_avgpool3s1p1(ctx, (const __half*)s.in, s.n, s.h, s.w, s.c, s.in_stride, (__half*)s.out,
                                 s.out_stride, st);
        break;
      case STEP_AVGPOOL_BN:
        rc = launch_avgpool_bn(ctx, (const __half*)s.in, s.n, s.h, s.w, s.c, s.pool_scale, s.pool_shift, s.out,
                               s.out_stride, s.out_f32, st);
        break;
      case STEP_GAP:
        rc = launch_global_avgpool(ctx, s.in, s.in_f32, s.n, s.h * s.w, s.c, s.in_stride, d_feat, st);
        break;
    }
    if (rc) return rc;
    if (ctx->profiling) {
      TISE_CUDA(ctx, cudaEventRecord(rec.b, st));
      ctx->prof.push_back(rec);
    }
  }
  if (progress_pending) TISE_CUDA(ctx, cudaEventRecord(ctx->progress_event, st));  // (a tap that ends before that layer)
  return TISE_OK;
}

}  // namespace tise
