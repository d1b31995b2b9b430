"""Inception-v3 forward of the reference's metrics.py on the jck kernels.

The reference evaluates IS / FID with `torchvision.models.inception_v3()` whose fc is replaced by Linear(2048, 100)
(metrics.py:46-52) and calls it in eval mode on 299 x 299 ImageNet-normalised images (metrics.py:80-93;
dcgan_trainer.py:203-207 does the 0.5x+0.5 / resize / normalise first).  Here the same network -- same state_dict keys,
so the reference's `./save/iception_v3/loss_bset.pt` loads unchanged -- runs as 94 implicit-GEMM launches of
`jck_conv_gemm` (tcgen05, bf16 operands, fp32 accumulation; eval-mode BatchNorm folded into the fp32 epilogue), 13 pooling
launches, one global average pool and the fc GEMM.  Activations are NHWC bf16 buffers that carry the zero border their
stride-1 consumer needs, so a kh x kw convolution is kh*kw shifted 2-D TMA boxes of the producer's buffer and the
concatenations of the Inception blocks are channel slices of one output buffer; only the five stride-2 convolutions go
through an explicit patch matrix (`jck_im2col`; the 3-channel stem's comes straight from the un-resized images, `jck_stem_patches`).

`K` is the kernel backend: `jck_generation_b200.ops` (the C ABI) -- there is no CPU path in the product.  The tests inject
`tests/incep_emul.py`, a torch restatement of the SAME primitives, to check this host graph on a machine without a GPU.
"""
import torch

BN_EPS = 1e-3          # torchvision BasicConv2d: nn.BatchNorm2d(out_channels, eps=0.001)
STRIDE2 = {"Conv2d_1a_3x3", "Mixed_6a.branch3x3", "Mixed_6a.branch3x3dbl_3", "Mixed_7a.branch3x3_2", "Mixed_7a.branch7x7x3_4"}
VALID = STRIDE2 | {"Conv2d_2a_3x3", "Conv2d_4a_3x3"}    # padding 0; every other convolution keeps the size
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def _ceil(v, m):
    return (v + m - 1) // m * m


def k_pitch(C):
    """columns per filter tap of a jck_conv_gemm weight matrix: C rounded up to the kernel's K chunk (32 for C <= 32, else 64)"""
    return 32 if C <= 32 else _ceil(C, 64)


class Buf:
    """NHWC bf16 activation buffer with a zero border of (py, px) pixels; channel pitch ld >= C.
    planes = 2 (split-precision mode): a second, identically laid out plane `lo` elements further on holds the low halves,
    value = hi + lo; R = rows (bordered pixels) of one plane, so in the GEMM's row space the low plane starts at row R."""

    def __init__(self, B, H, W, C, py=0, px=0, ld=None, device="cuda", dtype=torch.bfloat16, planes=1):
        self.B, self.H, self.W, self.C, self.py, self.px = B, H, W, C, py, px
        self.ld = ld or C
        self.Hb, self.Wb = H + 2 * py, W + 2 * px
        self.R = B * self.Hb * self.Wb
        self.planes = planes
        self.lo = self.R * self.ld if planes == 2 else 0
        self.t = torch.zeros(planes * self.R * self.ld, dtype=dtype, device=device)

    def geom(self, c_off=0):
        return [self.Hb, self.Wb, self.py, self.px, c_off]

    def _view(self, plane):
        n = self.R * self.ld
        return self.t[plane * n:(plane + 1) * n].view(self.B, self.Hb, self.Wb, self.ld)

    def interior(self):
        """[B, H, W, C] view of the logical tensor (tests); the HIGH plane in split mode."""
        return self._view(0)[:, self.py:self.py + self.H, self.px:self.px + self.W, :self.C]

    def value(self):
        """[B, H, W, C] fp32 values (hi + lo in split mode) -- tests"""
        v = self.interior().float()
        if self.planes == 2:
            v = v + self._view(1)[:, self.py:self.py + self.H, self.px:self.px + self.W, :self.C].float()
        return v


def split_bf16(w):
    """fp32 -> (hi, lo) bf16 with hi + lo = w to 16 mantissa bits"""
    hi = w.to(torch.bfloat16)
    return hi, (w - hi.float()).to(torch.bfloat16)


class _Conv:
    """One BasicConv2d: packed bf16 weight matrix + folded BatchNorm (fp32 scale, bias)."""

    def __init__(self, name, sd, device, dtype=torch.bfloat16, split=False):
        w = sd[name + ".conv.weight"].float()
        self.N, self.C, self.kh, self.kw = w.shape
        self.stride = 2 if name in STRIDE2 else 1
        self.pad = (0, 0) if name in VALID else ((self.kh - 1) // 2, (self.kw - 1) // 2)
        g, b = sd[name + ".bn.weight"].float(), sd[name + ".bn.bias"].float()
        m, v = sd[name + ".bn.running_mean"].float(), sd[name + ".bn.running_var"].float()
        scale = g / torch.sqrt(v + BN_EPS)
        self.scale = scale.contiguous().to(device)
        self.bias = (b - m * scale).contiguous().to(device)
        self.implicit = self.stride == 1 and self.C % 8 == 0
        wt = w.permute(0, 2, 3, 1).contiguous()                         # [N][kh][kw][C]
        if self.implicit:
            self.Cp = k_pitch(self.C)
            wm = torch.zeros(self.N, self.kh * self.kw, self.Cp)
            wm[:, :, :self.C] = wt.view(self.N, self.kh * self.kw, self.C)
        else:
            K = self.kh * self.kw * self.C
            self.Kp = _ceil(K, 8)
            wm = torch.zeros(self.N, k_pitch(self.Kp))
            wm[:, :K] = wt.reshape(self.N, K)
        if split:
            # [W_hi | W_lo | W_hi] along the tap axis: pairs with the activation taps {hi, hi, lo} (InceptionV3._conv)
            wm3 = wm.reshape(self.N, -1, wm.shape[-1])
            hi, lo = split_bf16(wm3)
            self.w = torch.cat([hi, lo, hi], dim=1).reshape(self.N, -1).contiguous().to(device)
        else:
            self.w = wm.reshape(self.N, -1).to(dtype).contiguous().to(device)


class InceptionV3:
    """forward(images) -> fp32 features [B, 100] (feature='logits', the reference's) or [B, 2048] ('pool3')."""

    def __init__(self, state_dict, feature="logits", device="cuda", K=None, act_dtype=torch.bfloat16, use_graph=True,
                 precision="bf16"):
        """precision: "bf16" -- activations and weights rounded to bf16 (fastest; 0.2-0.7 % per layer), or "split" -- every
        activation and weight kept as hi + lo bf16 pairs and every product formed as hi*hi + hi*lo + lo*hi on the same tcgen05
        kernel (3x the MMAs, 2x the activation bytes): fp32-grade features, the mode whose free-running features are pinned
        against torchvision fp32 (tests/test_gpu_incep.py)."""
        if K is None:
            from . import ops as K          # the C ABI; raises if libjck_b200.so is missing
            assert act_dtype == torch.bfloat16, "the kernels store activations in bf16 (fp32 is the emulator's check mode)"
        assert precision in ("bf16", "split"), precision
        assert precision == "bf16" or act_dtype == torch.bfloat16, "split precision is a pair of bf16 planes"
        self.split = precision == "split"
        self.planes = 2 if self.split else 1
        self.K, self.device, self.feature, self.dtype = K, torch.device(device), feature, act_dtype
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        names = sorted({k[:-len(".conv.weight")] for k in sd if k.endswith(".conv.weight") and not k.startswith("AuxLogits")})
        self.convs = {n: _Conv(n, sd, self.device, act_dtype, self.split) for n in names}
        self.fc_w = self.fc_b = None
        fcw = [k for k in sd if k.startswith("fc") and k.endswith("weight")]
        if feature == "logits":
            assert fcw, "state_dict has no fc layer"
            if self.split:
                hi, lo = split_bf16(sd[fcw[0]].float())
                self.fc_w = torch.cat([hi, lo, hi], dim=1).contiguous().to(self.device)      # [100][3 * 2048]
            else:
                self.fc_w = sd[fcw[0]].to(act_dtype).contiguous().to(self.device)          # [100][2048]
            self.fc_b = sd[fcw[0][:-len("weight")] + "bias"].float().contiguous().to(self.device)
        self._bufs = {}
        self._graphs = {}          # one CUDA graph of the 114 launches per input shape (the launches are 10-100 us each)
        self.use_graph = use_graph
        self.launches = 0

    # ---- buffers (allocated once per batch size; kernels write interiors only, so the borders stay zero) ----
    def _buf(self, key, B, H, W, C, py=0, px=0, ld=None):
        k = (key, B)
        if k not in self._bufs:
            self._bufs[k] = Buf(B, H, W, C, py, px, ld, self.device, self.dtype, self.planes)
        return self._bufs[k]

    def _scratch(self, key, numel, dtype):
        k = (key, dtype)
        if k not in self._bufs or self._bufs[k].numel() < numel:
            # growing a scratch buffer frees the old one, whose address is baked into every CUDA graph captured so far: drop
            # those graphs (they are re-captured on their next use) instead of letting a replay write into freed memory
            if k in self._bufs:
                self._graphs.clear()
            self._bufs[k] = torch.zeros(numel, dtype=dtype, device=self.device)
        return self._bufs[k]

    # ---- layers ----
    def _conv(self, name, src, dst, c_off=0):
        cv, K = self.convs[name], self.K
        B = src.B
        assert cv.C == src.C, (name, cv.C, src.C)
        Ho = (src.H + 2 * cv.pad[0] - cv.kh) // cv.stride + 1
        Wo = (src.W + 2 * cv.pad[1] - cv.kw) // cv.stride + 1
        assert (Ho, Wo) == (dst.H, dst.W) and c_off + cv.N <= dst.ld, (name, Ho, Wo, dst.H, dst.W)
        if cv.implicit:
            assert src.py >= cv.pad[0] and src.px >= cv.pad[1], (name, "source border too small")
            shifts = [(ky - cv.pad[0]) * src.Wb + (kx - cv.pad[1]) for ky in range(cv.kh) for kx in range(cv.kw)]
            rows = src.R
            taps, rows_a, tail = self._split_taps(shifts, rows, dst)
            geom = [rows, cv.N, cv.C, len(taps), src.Hb, src.Wb, src.py, src.px, Ho, Wo,
                    dst.Hb, dst.Wb, dst.py, dst.px, c_off, 1, 1, rows_a] + taps + tail
            K.conv_gemm(src.t, src.ld, cv.w, cv.scale, cv.bias, dst.t, dst.ld, geom)
            self.launches += 1
        else:
            M = B * Ho * Wo
            patches = self._scratch("patches", self.planes * M * cv.Kp, self.dtype)
            for pl in range(self.planes):           # split mode: the patch matrices of the two planes, stacked by rows
                K.im2col(src.t[pl * src.lo:], src.geom(), src.ld, patches[pl * M * cv.Kp:], B, src.H, src.W, src.C, cv.kh, cv.kw,
                         cv.stride, cv.stride, cv.pad[0], cv.pad[1], Ho, Wo, cv.Kp)
            taps, rows_a, tail = self._split_taps([0], M, dst)
            geom = [M, cv.N, cv.Kp, len(taps), Ho, Wo, 0, 0, Ho, Wo, dst.Hb, dst.Wb, dst.py, dst.px, c_off, 1, 1, rows_a] + taps + tail
            K.conv_gemm(patches, cv.Kp, cv.w, cv.scale, cv.bias, dst.t, dst.ld, geom)
            self.launches += 1 + self.planes
        return dst

    def _split_taps(self, shifts, rows, dst):
        """(tap shifts, rows of A, trailing geom) of a jck_conv_gemm call.  Split precision: the low plane of A starts `rows`
        rows behind the high one, so the three partial products hi*W_hi, hi*W_lo, lo*W_hi are the taps {s}, {s}, {rows + s}
        against the weight blocks [W_hi | W_lo | W_hi] (_Conv), and the output's own low plane lies dst.R rows behind."""
        if not self.split:
            return list(shifts), rows, []
        return list(shifts) + list(shifts) + [rows + s for s in shifts], 2 * rows, ([dst.R] if dst is not None else [])

    def _pool(self, src, dst, c_off, stride, pad, mode):
        Ho = (src.H + 2 * pad - 3) // stride + 1
        Wo = (src.W + 2 * pad - 3) // stride + 1
        assert (Ho, Wo) == (dst.H, dst.W)
        if self.split:
            self.K.pool3(src.t, src.geom(), src.ld, dst.t, dst.geom(c_off), dst.ld, src.B, src.H, src.W, src.C, stride, pad, Ho, Wo,
                         mode, x_lo=src.lo, out_lo=dst.lo)
        else:
            self.K.pool3(src.t, src.geom(), src.ld, dst.t, dst.geom(c_off), dst.ld, src.B, src.H, src.W, src.C, stride, pad, Ho, Wo, mode)
        self.launches += 1
        return dst

    def _block_a(self, p, x, pf):
        B, H, W = x.B, x.H, x.W
        out = self._buf(p + ".out", B, H, W, 224 + pf)
        self._conv(p + ".branch1x1", x, out, 0)
        t = self._conv(p + ".branch5x5_1", x, self._buf(p + ".b5", B, H, W, 48, 2, 2))
        self._conv(p + ".branch5x5_2", t, out, 64)
        t = self._conv(p + ".branch3x3dbl_1", x, self._buf(p + ".d1", B, H, W, 64, 1, 1))
        t = self._conv(p + ".branch3x3dbl_2", t, self._buf(p + ".d2", B, H, W, 96, 1, 1))
        self._conv(p + ".branch3x3dbl_3", t, out, 128)
        t = self._pool(x, self._buf(p + ".ap", B, H, W, x.C), 0, 1, 1, 1)
        self._conv(p + ".branch_pool", t, out, 224)
        return out

    def _block_b(self, p, x):
        B, H, W = x.B, x.H, x.W
        Ho, Wo = (H - 3) // 2 + 1, (W - 3) // 2 + 1
        out = self._buf(p + ".out", B, Ho, Wo, 384 + 96 + x.C)
        self._conv(p + ".branch3x3", x, out, 0)
        t = self._conv(p + ".branch3x3dbl_1", x, self._buf(p + ".d1", B, H, W, 64, 1, 1))
        t = self._conv(p + ".branch3x3dbl_2", t, self._buf(p + ".d2", B, H, W, 96))
        self._conv(p + ".branch3x3dbl_3", t, out, 384)
        self._pool(x, out, 480, 2, 0, 0)
        return out

    def _block_c(self, p, x, c7):
        B, H, W = x.B, x.H, x.W
        out = self._buf(p + ".out", B, H, W, 768)
        self._conv(p + ".branch1x1", x, out, 0)
        t = self._conv(p + ".branch7x7_1", x, self._buf(p + ".s1", B, H, W, c7, 0, 3))
        t = self._conv(p + ".branch7x7_2", t, self._buf(p + ".s2", B, H, W, c7, 3, 0))
        self._conv(p + ".branch7x7_3", t, out, 192)
        t = self._conv(p + ".branch7x7dbl_1", x, self._buf(p + ".d1", B, H, W, c7, 3, 0))
        t = self._conv(p + ".branch7x7dbl_2", t, self._buf(p + ".d2", B, H, W, c7, 0, 3))
        t = self._conv(p + ".branch7x7dbl_3", t, self._buf(p + ".d3", B, H, W, c7, 3, 0))
        t = self._conv(p + ".branch7x7dbl_4", t, self._buf(p + ".d4", B, H, W, c7, 0, 3))
        self._conv(p + ".branch7x7dbl_5", t, out, 384)
        t = self._pool(x, self._buf(p + ".ap", B, H, W, x.C), 0, 1, 1, 1)
        self._conv(p + ".branch_pool", t, out, 576)
        return out

    def _block_d(self, p, x):
        B, H, W = x.B, x.H, x.W
        Ho, Wo = (H - 3) // 2 + 1, (W - 3) // 2 + 1
        out = self._buf(p + ".out", B, Ho, Wo, 320 + 192 + x.C)
        t = self._conv(p + ".branch3x3_1", x, self._buf(p + ".t1", B, H, W, 192))
        self._conv(p + ".branch3x3_2", t, out, 0)
        t = self._conv(p + ".branch7x7x3_1", x, self._buf(p + ".s1", B, H, W, 192, 0, 3))
        t = self._conv(p + ".branch7x7x3_2", t, self._buf(p + ".s2", B, H, W, 192, 3, 0))
        t = self._conv(p + ".branch7x7x3_3", t, self._buf(p + ".s3", B, H, W, 192))
        self._conv(p + ".branch7x7x3_4", t, out, 320)
        self._pool(x, out, 512, 2, 0, 0)
        return out

    def _block_e(self, p, x):
        B, H, W = x.B, x.H, x.W
        out = self._buf(p + ".out", B, H, W, 2048)
        self._conv(p + ".branch1x1", x, out, 0)
        t = self._conv(p + ".branch3x3_1", x, self._buf(p + ".t1", B, H, W, 384, 1, 1))
        self._conv(p + ".branch3x3_2a", t, out, 320)
        self._conv(p + ".branch3x3_2b", t, out, 704)
        t = self._conv(p + ".branch3x3dbl_1", x, self._buf(p + ".d1", B, H, W, 448, 1, 1))
        t = self._conv(p + ".branch3x3dbl_2", t, self._buf(p + ".d2", B, H, W, 384, 1, 1))
        self._conv(p + ".branch3x3dbl_3a", t, out, 1088)
        self._conv(p + ".branch3x3dbl_3b", t, out, 1472)
        t = self._pool(x, self._buf(p + ".ap", B, H, W, x.C), 0, 1, 1, 1)
        self._conv(p + ".branch_pool", t, out, 1856)
        return out

    # ---- entry points ----
    def _stem(self, images, a, b, mean, std):
        """resize to 299 x 299 + normalise + Conv2d_1a_3x3 (3x3 stride 2): jck_stem_patches writes the stem's patch matrix
        straight from the input images, the GEMM applies the folded BatchNorm + ReLU"""
        B, C, Hi, Wi = images.shape
        assert C == 3, "Inception-v3 takes 3-channel images"
        cv = self.convs["Conv2d_1a_3x3"]
        M = B * 149 * 149
        patches = self._scratch("patches", self.planes * M * 32, self.dtype)
        if self.split:
            self.K.stem_patches(images, patches, B, Hi, Wi, 299, 299, a, b, mean, std, patches_lo=M * 32)
        else:
            self.K.stem_patches(images, patches, B, Hi, Wi, 299, 299, a, b, mean, std)
        dst = self._buf("c1a", B, 149, 149, 32)
        taps, rows_a, tail = self._split_taps([0], M, dst)
        geom = [M, cv.N, 32, len(taps), 149, 149, 0, 0, 149, 149, dst.Hb, dst.Wb, 0, 0, 0, 1, 1, rows_a] + taps + tail
        self.K.conv_gemm(patches, 32, cv.w, cv.scale, cv.bias, dst.t, dst.ld, geom)
        self.launches += 2
        return dst

    def _eager(self, images, a, b, mean, std):
        return self._trunk(self._stem(images, a, b, mean, std))

    def _run(self, images, a, b, mean, std):
        images = images.float().contiguous()
        if not (self.use_graph and images.is_cuda):
            return self._eager(images, a, b, mean, std)
        key = (tuple(images.shape), a, b)
        if key not in self._graphs:
            static_in = images.clone()
            self._eager(static_in, a, b, mean, std)            # allocates the buffers, sets kernel attributes
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._eager(static_in, a, b, mean, std)
            self._graphs[key] = (graph, static_in, out)
        graph, static_in, out = self._graphs[key]
        static_in.copy_(images)
        graph.replay()
        return out.clone()

    def forward(self, images):
        """images: NCHW fp32 on the device, already ImageNet-normalised (what metrics.py's loaders hold); any size is resized
        to 299 x 299 bilinearly (identity for 299 x 299 inputs)."""
        return self._run(images, 1.0, 0.0, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0))

    def forward_generated(self, fake):
        """fake: the generator's tanh output in [-1, 1], NCHW fp32.  Fuses dcgan_trainer.py:203-207 (0.5x + 0.5, F.resize to
        299 x 299, ImageNet normalise) into the stem's patch kernel."""
        return self._run(fake, 0.5, 0.5, IMAGENET_MEAN, IMAGENET_STD)

    def _stages(self):
        """the trunk after the stem as (output buffer key, function of the previous stage's buffer)"""
        return [
            ("c2a", lambda x: self._conv("Conv2d_2a_3x3", x, self._buf("c2a", x.B, 147, 147, 32, 1, 1))),
            ("c2b", lambda x: self._conv("Conv2d_2b_3x3", x, self._buf("c2b", x.B, 147, 147, 64))),
            ("p1", lambda x: self._pool(x, self._buf("p1", x.B, 73, 73, 64), 0, 2, 0, 0)),
            ("c3b", lambda x: self._conv("Conv2d_3b_1x1", x, self._buf("c3b", x.B, 73, 73, 80))),
            ("c4a", lambda x: self._conv("Conv2d_4a_3x3", x, self._buf("c4a", x.B, 71, 71, 192))),
            ("p2", lambda x: self._pool(x, self._buf("p2", x.B, 35, 35, 192), 0, 2, 0, 0)),
            ("Mixed_5b.out", lambda x: self._block_a("Mixed_5b", x, 32)),
            ("Mixed_5c.out", lambda x: self._block_a("Mixed_5c", x, 64)),
            ("Mixed_5d.out", lambda x: self._block_a("Mixed_5d", x, 64)),
            ("Mixed_6a.out", lambda x: self._block_b("Mixed_6a", x)),
            ("Mixed_6b.out", lambda x: self._block_c("Mixed_6b", x, 128)),
            ("Mixed_6c.out", lambda x: self._block_c("Mixed_6c", x, 160)),
            ("Mixed_6d.out", lambda x: self._block_c("Mixed_6d", x, 160)),
            ("Mixed_6e.out", lambda x: self._block_c("Mixed_6e", x, 192)),
            ("Mixed_7a.out", lambda x: self._block_d("Mixed_7a", x)),
            ("Mixed_7b.out", lambda x: self._block_e("Mixed_7b", x)),
            ("Mixed_7c.out", lambda x: self._block_e("Mixed_7c", x)),
        ]

    def _trunk(self, x):
        B = x.B
        for _, stage in self._stages():
            x = stage(x)
        return self._head(x)

    def _head(self, x):
        B = x.B
        pooled = torch.empty(B, 2048, dtype=torch.float32, device=self.device)
        pooled_bf = self._scratch("pooled_bf", self.planes * B * 2048, self.dtype)
        if self.split:
            self.K.global_avgpool(x.t, pooled, pooled_bf, B, x.H * x.W, 2048, x_lo=x.lo, out_lo=B * 2048)
        else:
            self.K.global_avgpool(x.t, pooled, pooled_bf, B, x.H * x.W, 2048)
        self.launches += 1
        if self.feature == "pool3":
            return pooled
        n = self.fc_w.shape[0]
        logits = torch.empty(B, n, dtype=torch.float32, device=self.device)
        taps, rows_a, _ = self._split_taps([0], B, None)
        geom = [B, n, 2048, len(taps), 1, 1, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 0, rows_a] + taps
        self.K.conv_gemm(pooled_bf, 2048, self.fc_w, None, self.fc_b, logits, n, geom)
        self.launches += 1
        return logits
