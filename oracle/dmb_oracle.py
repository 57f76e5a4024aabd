"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the cost-volume hot path.

A functional, CPU, fp32 restatement of the reference's algorithm for every function on the
hot path.  It exists so that the CUDA kernels can be checked on a
box where `/root/reference` is absent.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import it, and only as the checker
/ the reported CPU baseline -- the product path (`densematchingbenchmark_b200`) never does.

Pinning: every function that has a counterpart under `/root/reference` is pinned against
outputs of the reference itself (`oracle/make_golden.py` imports the reference through
`oracle/ref_import.py`, writes `tests/golden/*.pt`; `tests/test_oracle_golden.py` replays
them).  The functions with NO counterpart in the reference snapshot -- `gwc_volume`,
`sga`, `lga` -- are **parity unpinned**: they restate the published
GwcNet / GANet definitions and are the contract for the kernels by themselves.
`correlation1d_cost` restates the published definition of the un-vendored `spatial_correlation_sampler`
dependency -- **parity unpinned** as well.  `spn_scan*` restates `dmb/ops/spn/src/gaterecurrent2dnoind_kernel.cu`:
the CPU restatement cannot be pinned in this container (CUDA only, no CPU path), but on the GPU box it IS pinned --
`tests/test_gpu_spn_ref.py` checks it (and csrc/scans.cu) against the reference kernel itself, compiled from the
reference tree into the test-only `oracle/_ref/libspn_ref.so` by `oracle/build_ref.py`.

Everything here is deliberately simple: explicit index arithmetic and torch CPU tensor ops
in float32 (the reference's dtype), no modules, weights addressed through reference
state-dict key names.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm3d default, layers/basic_layers.py:74


# --------------------------------------------------------------------------------------
# stereo focal loss (the consumer of the raw cost volumes in AcfNet training)
# --------------------------------------------------------------------------------------
def laplace_disp2prob(gt, max_disp, variance=1.0, start_disp=0, dilation=1, disp_sample=None):
    """losses/utils/disp2prob.py:29-173 (Disp2Prob.getProb + LaplaceDisp2Prob): ground-truth disparity map
    [B,1,H,W] -> probability volume [B,n,H,W]: softmax over the samples of -|d - gt| / variance, zeroed where gt
    is outside (start, start + max_disp - 1) (the INNER mask, with `end = start + max_disp - 1`, :60,:126-128),
    plus eps = 1e-40 (:63,:137)."""
    B, _, H, W = gt.shape
    end = start_disp + max_disp - 1
    if disp_sample is None:
        n = (max_disp + dilation - 1) // dilation
        disp_sample = torch.linspace(start_disp, end, n).view(1, n, 1, 1).expand(B, n, H, W)
    mask = ((gt > start_disp) & (gt < end)).to(gt.dtype)
    gt = gt * mask
    cost = -torch.abs(disp_sample - gt) / variance
    return F.softmax(cost, dim=1) * mask + 1e-40


def stereo_focal_loss(est_cost, gt_disp, variance, max_disp, start_disp=0, dilation=1, focal_coefficient=0.0,
                      sparse=False, disp_sample=None):
    """StereoFocalLoss.loss_per_level (losses/stereo_focal_loss.py:63-101) for one cost volume [B,D,H,W]:
    the ground truth is rescaled (and average- / max-pooled) to the cost's resolution (:66-73), masked to
    (start, start + int(max_disp / scale)) (:78-81), turned into a Laplace distribution over the disparity samples,
    and  loss = -sum(gtProb * log_softmax(cost) * (1 - gtProb)^(-coefficient) * mask) / max(#valid, 1)  (:95-99)."""
    B, D, H, W = est_cost.shape
    gt = gt_disp.clone()
    scale = 1.0
    if gt_disp.shape[-2] != H or gt_disp.shape[-1] != W:
        scale = gt_disp.shape[-1] / (W * 1.0)
        gt = gt / scale
        gt = (F.adaptive_max_pool2d if sparse else F.adaptive_avg_pool2d)(gt, (H, W))
    lower = start_disp
    upper = lower + int(max_disp / scale)
    mask = ((gt > lower) & (gt < upper)).to(gt.dtype)
    if float(mask.sum()) < 1.0:
        gt_prob = torch.zeros_like(est_cost)
    else:
        gt_prob = laplace_disp2prob(gt * mask, int(max_disp / scale), variance, start_disp, dilation, disp_sample)
    valid = float(mask.sum())
    if valid < 1.0:
        valid = 1.0
    log_prob = F.log_softmax(est_cost, dim=1)
    weight = (1.0 - gt_prob).pow(-focal_coefficient)
    return -((gt_prob * log_prob) * weight * mask).sum() / valid


# --------------------------------------------------------------------------------------
# disparity sampling
# --------------------------------------------------------------------------------------
def disp_count(max_disp, dilation=1):
    """cat_fms.py:28 -- number of disparity samples."""
    return (max_disp + dilation - 1) // dilation


def disp_indices(max_disp, start_disp=0, dilation=1):
    """cat_fms.py:27-35 -- `linspace(start, start+max-1, n)` then `int()` (truncation toward
    zero of the float32 linspace value).  Returns a python list of ints."""
    n = disp_count(max_disp, dilation)
    vals = torch.linspace(start_disp, start_disp + max_disp - 1, n)
    return [int(v) for v in vals]


def disp_samples(max_disp, start_disp=0, dilation=1):
    """soft_argmin.py:40-42 / faster_soft_argmin.py:41-43 -- the float disparity samples."""
    n = disp_count(max_disp, dilation)
    return torch.linspace(start_disp, start_disp + max_disp - 1, n)


# --------------------------------------------------------------------------------------
# raw cost volumes
# --------------------------------------------------------------------------------------
def _shift_pair(left, right, d):
    """Return (left part, right part, valid x-slice) for integer disparity d
    (cat_fms.py:36-44): d>0 -> x in [d,W): L[x], R[x-d];  d<0 -> x in [0,W+d): L[x], R[x-d]."""
    W = left.shape[-1]
    if d > 0:
        return left[..., d:], right[..., :W - d], slice(d, W)
    if d == 0:
        return left, right, slice(0, W)
    return left[..., :d], right[..., -d:], slice(0, W + d)


def cat_volume(left, right, max_disp=192, start_disp=0, dilation=1):
    """cat_fms (cost_processors/utils/cat_fms.py:7-48) -> [B,2C,D,H,W] float32."""
    B, C, H, W = left.shape
    idx = disp_indices(max_disp, start_disp, dilation)
    out = torch.zeros(B, 2 * C, len(idx), H, W, dtype=torch.float32, device=left.device)
    for k, d in enumerate(idx):
        if abs(d) >= W:
            continue  # empty slice in the reference
        l, r, xs = _shift_pair(left, right, d)
        out[:, :C, k, :, xs] = l
        out[:, C:, k, :, xs] = r
    return out


def dif_volume(left, right, max_disp=192, start_disp=0, dilation=1):
    """dif_fms (cost_processors/utils/dif_fms.py:7-46) -> [B,C,D,H,W] float32."""
    B, C, H, W = left.shape
    idx = disp_indices(max_disp, start_disp, dilation)
    out = torch.zeros(B, C, len(idx), H, W, dtype=torch.float32)
    for k, d in enumerate(idx):
        if abs(d) >= W:
            continue
        l, r, xs = _shift_pair(left, right, d)
        out[:, :, k, :, xs] = l - r
    return out


def warp_volume(target, disp_sample):
    """inverse_warp_3d(target expanded over D, -disp_sample) (layers/inverse_warp_3d.py:4-52)
    as called by fast_cat_fms / fast_dif_fms (cat_fms.py:74, dif_fms.py:75).

    The reference normalises with (size-1) (align_corners=True convention, :40-42) but calls
    F.grid_sample with the default align_corners=False (:50).  Un-normalising with the
    align_corners=False rule gives the sample position
        ix = (x - disp) * W/(W-1) - 0.5,  iy = y * H/(H-1) - 0.5,  iz = k * D/(D-1) - 0.5
    followed by trilinear interpolation with zero padding.  Because the source volume is the
    target feature map replicated along D, the z interpolation only scales by the in-range
    z weight.  Returns [B,C,D,H,W]."""
    B, C, H, W = target.shape
    D = disp_sample.shape[1]
    f32 = torch.float32
    gx = torch.arange(W, dtype=f32).view(1, 1, 1, W) - disp_sample.to(f32)
    gy = torch.arange(H, dtype=f32).view(1, 1, H, 1).expand(B, D, H, W)
    gz = torch.arange(D, dtype=f32).view(1, D, 1, 1).expand(B, D, H, W)
    # reference arithmetic order (:40-42) then grid_sample's unnormalise ((g+1)*size-1)/2
    nx = gx / (W - 1) * 2 - 1
    ny = gy / (H - 1) * 2 - 1
    nz = gz / (D - 1) * 2 - 1
    ix = ((nx + 1) * W - 1) / 2
    iy = ((ny + 1) * H - 1) / 2
    iz = ((nz + 1) * D - 1) / 2
    x0 = torch.floor(ix); y0 = torch.floor(iy); z0 = torch.floor(iz)
    out = torch.zeros(B, C, D, H, W, dtype=f32)
    bidx = torch.arange(B).view(B, 1, 1, 1).expand(B, D, H, W)
    for dz in (0, 1):
        zz = z0 + dz
        wz = (1 - (iz - z0)) if dz == 0 else (iz - z0)
        okz = (zz >= 0) & (zz <= D - 1)
        for dy in (0, 1):
            yy = y0 + dy
            wy = (1 - (iy - y0)) if dy == 0 else (iy - y0)
            oky = (yy >= 0) & (yy <= H - 1)
            for dx in (0, 1):
                xx = x0 + dx
                wx = (1 - (ix - x0)) if dx == 0 else (ix - x0)
                okx = (xx >= 0) & (xx <= W - 1)
                ok = (okx & oky & okz).to(f32)
                yi = yy.clamp(0, H - 1).long(); xi = xx.clamp(0, W - 1).long()
                val = target[bidx, :, yi, xi]            # [B,D,H,W,C]
                out += (val * (wx * wy * wz * ok).unsqueeze(-1)).permute(0, 4, 1, 2, 3)
    return out


def _default_disp_sample(B, H, W, max_disp, start_disp, dilation):
    s = disp_samples(max_disp, start_disp, dilation)
    return s.view(1, -1, 1, 1).expand(B, s.numel(), H, W).float()


def fast_cat_volume(left, right, max_disp=192, start_disp=0, dilation=1, disp_sample=None):
    """fast_cat_fms (cat_fms.py:51-82): warp target, mask reference where warped target <= 0."""
    B, C, H, W = left.shape
    if disp_sample is None:
        disp_sample = _default_disp_sample(B, H, W, max_disp, start_disp, dilation)
    tgt = warp_volume(right, disp_sample)
    ref = left.unsqueeze(2) * (tgt > 0).float()
    return torch.cat((ref, tgt), dim=1)


def fast_dif_volume(left, right, max_disp=192, start_disp=0, dilation=1, disp_sample=None,
                    normalize=False, p=1.0):
    """fast_dif_fms (dif_fms.py:49-86)."""
    B, C, H, W = left.shape
    if disp_sample is None:
        disp_sample = _default_disp_sample(B, H, W, max_disp, start_disp, dilation)
    tgt = warp_volume(right, disp_sample)
    ref = left.unsqueeze(2) * (tgt > 0).float()
    dif = ref - tgt
    if normalize:
        dif = torch.norm(dif, p=p, dim=1, keepdim=False)
    return dif


def gwc_volume(left, right, num_groups, max_disp=192, start_disp=0, dilation=1):
    """PARITY UNPINNED (no reference code).  Group-wise correlation of
    GwcNet (Guo et al., CVPR 2019, eq. 3): cost[b,g,k,y,x] = mean over the C/G channels of
    group g of L[b,c,y,x] * R[b,c,y,x-d_k], zero where the shifted pixel is outside the image.
    Disparities enumerated exactly like cat_fms (cat_fms.py:27-35)."""
    B, C, H, W = left.shape
    assert C % num_groups == 0
    cpg = C // num_groups
    idx = disp_indices(max_disp, start_disp, dilation)
    out = torch.zeros(B, num_groups, len(idx), H, W, dtype=torch.float32)
    for k, d in enumerate(idx):
        if abs(d) >= W:
            continue
        l, r, xs = _shift_pair(left, right, d)
        prod = (l * r).view(B, num_groups, cpg, H, l.shape[-1])
        out[:, :, k, :, xs] = prod.mean(dim=2)
    return out


def correlation1d_cost(left, right, max_disp=192, negative_slope=0.1):
    """PARITY UNPINNED.  correlation1d_cost (cost_processors/utils/correlation1d_cost.py:7-27).  The reference calls
    `spatial_correlation_sampler.SpatialCorrelationSampler(patch_size=(1, 2*max_disp-1), kernel_size=1, stride=1,
    padding=0, dilation_patch=1)` -- a third-party CUDA extension that is neither vendored nor version-pinned by the
    reference (INSTALL.md:60 names the repository ClementPinard/Pytorch-Correlation-extension) and cannot be imported
    here.  Its published definition, restated: out[b, ph, pw, y, x] = sum_c in1[b,c,y,x] * in2[b,c,y+ph-PH//2,x+pw-PW//2]
    with zero padding of in2 and no normalisation.  With PH = 1, PW = 2*max_disp-1 the reference squeezes ph, keeps
    pw = 0 .. max_disp-1 (shifts -(max_disp-1) .. 0, :19-22) and applies leaky ReLU(0.1) (:24):
        out[b, j, y, x] = lrelu( sum_c L[b,c,y,x] * R[b,c,y,x - (max_disp-1-j)] )."""
    B, C, H, W = left.shape
    out = torch.zeros(B, max_disp, H, W, dtype=torch.float32)
    for j in range(max_disp):
        d = max_disp - 1 - j
        if d >= W:
            continue
        l, r, xs = _shift_pair(left, right, d)
        out[:, j, :, xs] = (l * r).sum(dim=1)
    return F.leaky_relu(out, negative_slope)


# --------------------------------------------------------------------------------------
# 3-D conv building blocks, addressed by reference state-dict keys
# --------------------------------------------------------------------------------------
BN_MOMENTUM = 0.1  # nn.BatchNorm3d default


def _bn(sd, key, x, train=None):
    """BatchNorm3d (nn.BatchNorm3d, basic_layers.py:74).  `train=None`: eval mode, running statistics.
    `train` = dict: training mode -- biased batch statistics over (B,D,H,W) normalise the activations and
    the running statistics the module would hold afterwards (momentum 0.1, unbiased variance,
    num_batches_tracked + 1) are recorded in train['running']."""
    shape = (1, -1, 1, 1, 1)
    w = sd[key + ".weight"].view(shape); b = sd[key + ".bias"].view(shape)
    if train is not None:
        dims = (0, 2, 3, 4)
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        n = x.numel() // x.shape[1]
        run = train.setdefault("running", {})
        with torch.no_grad():
            run[key + ".running_mean"] = (1 - BN_MOMENTUM) * sd[key + ".running_mean"] + BN_MOMENTUM * mean
            run[key + ".running_var"] = ((1 - BN_MOMENTUM) * sd[key + ".running_var"]
                                         + BN_MOMENTUM * var * (n / max(n - 1, 1)))
            run[key + ".num_batches_tracked"] = sd[key + ".num_batches_tracked"] + 1
        return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS) * w + b
    m = sd[key + ".running_mean"].view(shape); v = sd[key + ".running_var"].view(shape)
    return (x - m) / torch.sqrt(v + BN_EPS) * w + b


def conv_unit(sd, key, x, stride=1, relu=False, transposed=False, batch_norm=True, train=None):
    """One `nn.Sequential(Conv3d|ConvTranspose3d, [BatchNorm3d], [ReLU])` of
    layers/basic_layers.py:68-216: `<key>.0` is the conv, `<key>.1` the BN.  3x3x3, padding 1,
    (transposed: output_padding 1 as in hourglass.py:53-60)."""
    w = sd[key + ".0.weight"]
    b = sd.get(key + ".0.bias")
    if transposed:
        y = F.conv_transpose3d(x, w, b, stride=stride, padding=1, output_padding=stride - 1)
    else:
        y = F.conv3d(x, w, b, stride=stride, padding=1)
    if batch_norm:
        y = _bn(sd, key + ".1", y, train)
    return F.relu(y) if relu else y


def hourglass(sd, key, x, presqu=None, postsqu=None, batch_norm=True, train=None):
    """Hourglass.forward (cost_processors/utils/hourglass.py:62-86)."""
    bn = dict(batch_norm=batch_norm, train=train)
    out = conv_unit(sd, key + ".conv1", x, stride=2, relu=True, **bn)
    pre = conv_unit(sd, key + ".conv2", out, **bn)
    pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
    out = conv_unit(sd, key + ".conv3", pre, stride=2, relu=True, **bn)
    out = conv_unit(sd, key + ".conv4", out, relu=True, **bn)
    up = conv_unit(sd, key + ".conv5", out, stride=2, transposed=True, **bn)
    post = F.relu(up + (presqu if presqu is not None else pre))
    out = conv_unit(sd, key + ".conv6", post, stride=2, transposed=True, **bn)
    return out, pre, post


def _classif(sd, key, x, batch_norm=True, train=None):
    y = conv_unit(sd, key + ".0", x, relu=True, batch_norm=batch_norm, train=train)
    return F.conv3d(y, sd[key + ".1.weight"], sd.get(key + ".1.bias"), padding=1)


def psm_trunk(sd, raw_cost, prefix="", batch_norm=True, train=None):
    """PSMAggregator.forward up to the three low-res costs (aggregators/PSMNet.py:55-72);
    AcfAggregator shares it (aggregators/AcfNet.py:62-76).  Returns (cost1, cost2, cost3),
    each [B,1,D4,H4,W4]."""
    p = prefix
    bn = dict(batch_norm=batch_norm, train=train)
    c0 = conv_unit(sd, p + "dres0.0", raw_cost, relu=True, **bn)
    c0 = conv_unit(sd, p + "dres0.1", c0, relu=True, **bn)
    t = conv_unit(sd, p + "dres1.0", c0, relu=True, **bn)
    c0 = conv_unit(sd, p + "dres1.1", t, **bn) + c0
    o1, pre1, post1 = hourglass(sd, p + "dres2", c0, None, None, batch_norm, train)
    o1 = o1 + c0
    o2, pre2, post2 = hourglass(sd, p + "dres3", o1, pre1, post1, batch_norm, train)
    o2 = o2 + c0
    o3, _, _ = hourglass(sd, p + "dres4", o2, pre2, post2, batch_norm, train)
    o3 = o3 + c0
    cost1 = _classif(sd, p + "classif1", o1, batch_norm, train)
    cost2 = _classif(sd, p + "classif2", o2, batch_norm, train) + cost1
    cost3 = _classif(sd, p + "classif3", o3, batch_norm, train) + cost2
    return cost1, cost2, cost3


def trilinear_up(cost, out_dhw):
    """F.interpolate(mode='trilinear', align_corners=True) (aggregators/PSMNet.py:75-88),
    written out: src = dst * (in-1)/(out-1), linear blend of the two neighbours per axis."""
    x = cost
    for axis, n_out in zip((2, 3, 4), out_dhw):
        n_in = x.shape[axis]
        if n_in == n_out:
            continue
        scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        pos = torch.arange(n_out, dtype=torch.float32) * scale
        i0 = pos.floor().long().clamp(max=n_in - 1)
        i1 = (i0 + 1).clamp(max=n_in - 1)
        w1 = pos - i0.float()
        shape = [1] * x.dim(); shape[axis] = n_out
        w1 = w1.view(shape)
        x = x.index_select(axis, i0) * (1 - w1) + x.index_select(axis, i1) * w1
    return x


def psm_aggregator(sd, raw_cost, max_disp, prefix="", batch_norm=True, train=None):
    """PSMAggregator.forward (aggregators/PSMNet.py:55-95) -> [cost3, cost2, cost1]."""
    B, C, D, H, W = raw_cost.shape
    c1, c2, c3 = psm_trunk(sd, raw_cost, prefix, batch_norm, train)
    size = [max_disp, H * 4, W * 4]
    up = [F.interpolate(c, size, mode="trilinear", align_corners=True).squeeze(1) for c in (c3, c2, c1)]
    return up


def acf_aggregator(sd, raw_cost, max_disp, prefix="", batch_norm=True, train=None):
    """AcfAggregator.forward (aggregators/AcfNet.py:59-89): same trunk, learned
    ConvTranspose3d(1,1,8,4,2) upsampling (:55-57,81-83)."""
    c1, c2, c3 = psm_trunk(sd, raw_cost, prefix, batch_norm, train)
    outs = []
    for c, name in ((c3, "deconv3"), (c2, "deconv2"), (c1, "deconv1")):
        outs.append(F.conv_transpose3d(c, sd[prefix + name + ".weight"], None, stride=4, padding=2).squeeze(1))
    return outs


# --------------------------------------------------------------------------------------
# confidence measurement network
# --------------------------------------------------------------------------------------
def conf_head(sd, prefix, cost):
    """ConfHead.forward in eval mode (cmn/cmn.py:26-37): conv_bn_relu(in_planes, in_planes // 3, 3x3, bias=False)
    (layers/basic_layers.py:103-119) then Conv2d(in_planes // 3, 1, 1x1, bias=False); `prefix` ends with 'conf_net.'."""
    y = F.conv2d(cost, sd[prefix + "0.0.weight"], None, padding=1)
    y = F.batch_norm(y, sd[prefix + "0.1.running_mean"], sd[prefix + "0.1.running_var"], sd[prefix + "0.1.weight"],
                     sd[prefix + "0.1.bias"], False, 0.0, BN_EPS)
    return F.conv2d(F.relu(y), sd[prefix + "1.weight"])


def cmn_eval(sd, costs, alpha, beta, prefix="conf_heads."):
    """Cmn.get_confidence (cmn/cmn.py:57-70): confidence = sigmoid(conf cost), variance = alpha * (1 - conf) + beta."""
    confs = [torch.sigmoid(conf_head(sd, "%s%d.conf_net." % (prefix, i), c)) for i, c in enumerate(costs)]
    return confs, [alpha * (1 - c) + beta for c in confs]


# --------------------------------------------------------------------------------------
# disparity regression
# --------------------------------------------------------------------------------------
def soft_argmin(cost, max_disp, start_disp=0, dilation=1, alpha=1.0, normalize=True, disp_sample=None):
    """SoftArgmin.forward (disp_predictors/soft_argmin.py:44-75) and FasterSoftArgmin.forward
    (faster_soft_argmin.py:51-75; its frozen Conv3d(1,1,(D,1,1)) holds the same linspace,
    :41-49): softmax over D of cost*alpha, expectation of the disparity samples."""
    if cost.dim() != 4:
        raise ValueError("expected 4D input (got {}D input)".format(cost.dim()))
    c = cost * alpha
    p = torch.softmax(c, dim=1) if normalize else c
    if disp_sample is None:
        s = disp_samples(max_disp, start_disp, dilation)
        assert s.numel() == cost.shape[1]
        disp_sample = s.view(1, -1, 1, 1).to(cost.device)
    return (p * disp_sample).sum(dim=1, keepdim=True)


def local_soft_argmin(cost, max_disp, radius, start_disp=0, dilation=1, radius_dilation=1, alpha=1.0):
    """LocalSoftArgmin.forward (disp_predictors/local_soft_argmin.py:47-105): argmax over D,
    window of 2*radius+1 indices spaced radius_dilation, out-of-range taps get logit
    -10000*alpha but keep their *clamped* index as disparity value (:75-89)."""
    B, D, H, W = cost.shape
    assert D == disp_count(max_disp, dilation)
    best = cost.argmax(dim=1, keepdim=True)
    offs = torch.linspace(-radius * radius_dilation, radius * radius_dilation, 2 * radius + 1).long()
    idx = best + offs.view(1, -1, 1, 1)
    inside = ((idx >= 0) & (idx <= D - 1)).float()
    idx = idx.clamp(0, D - 1)
    g = torch.gather(cost, 1, idx) * alpha
    logits = g * inside + (1 - inside) * (-10000.0 * alpha)
    p = torch.softmax(logits, dim=1)
    disp = start_disp + idx.float() * dilation
    return (p * disp).sum(dim=1, keepdim=True)


# --------------------------------------------------------------------------------------
# dmb.ops: SPN 3-neighbour gated scan
# --------------------------------------------------------------------------------------
def spn_scan(X, G1, G2, G3, horizontal, reverse):
    """PARITY UNPINNED (CUDA-only reference op).  Restates forward_one_col_left_right &
    siblings (dmb/ops/spn/src/gaterecurrent2dnoind_kernel.cu:130-285):
        h(p) = (1-g1-g2-g3)*x(p) + g1*h(p - s + perp(-1)) + g2*h(p - s) + g3*h(p - s + perp(+1))
    where s is one step along the scan direction and perp(+-1) a step across it.  A gate whose
    neighbour lies outside the image reads 0 (get_gate_sf, :79-97); otherwise the gate is read
    at the *current* pixel (the later of the two in scan order, get_gate_idx_sf :10-65);
    out-of-image h reads 0 (get_data_sf :67-75)."""
    N, C, H, W = X.shape
    out = torch.zeros_like(X)
    if horizontal:
        steps = range(W - 1, -1, -1) if reverse else range(W)
        prev_off = 1 if reverse else -1
        for t in steps:
            tp = t + prev_off
            x = X[..., t]
            if 0 <= tp < W:
                hp = out[..., tp]                                   # [N,C,H]
                up = F.pad(hp, (1, 0))[..., :H]                     # h[y-1]
                dn = F.pad(hp, (0, 1))[..., 1:]                     # h[y+1]
                g1 = G1[..., t].clone(); g1[..., 0] = 0            # neighbour y-1 outside at y=0
                g2 = G2[..., t]
                g3 = G3[..., t].clone(); g3[..., H - 1] = 0
            else:
                z = torch.zeros_like(x)
                up = hp = dn = g1 = g2 = g3 = z
            out[..., t] = (1 - g1 - g2 - g3) * x + g1 * up + g2 * hp + g3 * dn
    else:
        steps = range(H - 1, -1, -1) if reverse else range(H)
        prev_off = 1 if reverse else -1
        for t in steps:
            tp = t + prev_off
            x = X[:, :, t, :]
            if 0 <= tp < H:
                hp = out[:, :, tp, :]                               # [N,C,W]
                lf = F.pad(hp, (1, 0))[..., :W]                     # h[x-1]
                rt = F.pad(hp, (0, 1))[..., 1:]                     # h[x+1]
                g1 = G1[:, :, t, :].clone(); g1[..., 0] = 0
                g2 = G2[:, :, t, :]
                g3 = G3[:, :, t, :].clone(); g3[..., W - 1] = 0
            else:
                z = torch.zeros_like(x)
                lf = hp = rt = g1 = g2 = g3 = z
            out[:, :, t, :] = (1 - g1 - g2 - g3) * x + g1 * lf + g2 * hp + g3 * rt
    return out


def spn_scan_backward(X, G1, G2, G3, Hout, grad_out, horizontal, reverse):
    """PARITY UNPINNED.  Gradients of `spn_scan` obtained by differentiating the restated
    forward recurrence with autograd (the hand-written backward kernels,
    gaterecurrent2dnoind_kernel.cu:288-532, implement the same adjoint recurrence; note the
    reference leaves gate-grads untouched (=0) where the neighbour is outside, set_gate_sf
    :99-121, which is what masking the gate to 0 in the forward yields)."""
    Xr = X.clone().requires_grad_(True)
    Gs = [g.clone().requires_grad_(True) for g in (G1, G2, G3)]
    out = _spn_autograd(Xr, Gs[0], Gs[1], Gs[2], horizontal, reverse)
    out.backward(grad_out)
    # a scan of length 1 never touches the gates: autograd leaves their .grad unset (= zero)
    return tuple(t.grad if t.grad is not None else torch.zeros_like(t) for t in [Xr] + Gs)


def _spn_autograd(X, G1, G2, G3, horizontal, reverse):
    """Out-of-place twin of spn_scan for autograd."""
    if not horizontal:  # scan over H == horizontal scan of the transposed problem
        o = _spn_autograd(X.transpose(2, 3), G1.transpose(2, 3), G2.transpose(2, 3), G3.transpose(2, 3), True, reverse)
        return o.transpose(2, 3)
    N, C, H, W = X.shape
    cols = [None] * W
    order = range(W - 1, -1, -1) if reverse else range(W)
    prev_off = 1 if reverse else -1
    m1 = torch.ones(H); m1[0] = 0
    m3 = torch.ones(H); m3[H - 1] = 0
    for t in order:
        tp = t + prev_off
        x = X[..., t]
        if 0 <= tp < W:
            hp = cols[tp]
            up = F.pad(hp, (1, 0))[..., :H]
            dn = F.pad(hp, (0, 1))[..., 1:]
            g1 = G1[..., t] * m1; g2 = G2[..., t]; g3 = G3[..., t] * m3
            cols[t] = (1 - g1 - g2 - g3) * x + g1 * up + g2 * hp + g3 * dn
        else:
            cols[t] = x * 1.0
    return torch.stack(cols, dim=-1)


# --------------------------------------------------------------------------------------
# GANet aggregation layers (no reference code)
# --------------------------------------------------------------------------------------
def sga(x, guidance):
    """PARITY UNPINNED.  Semi-global aggregation layer of GANet (Zhang et al., CVPR 2019,
    eq. 5).  x: [B,C,D,H,W]; guidance: [B,4,5,C,H,W]... stored as [B, 4*5*C, H, W] ordered
    (direction, tap, channel) -- directions: 0 left->right, 1 right->left, 2 top->bottom,
    3 bottom->top; taps w0..w4, L1-normalised over the 5 taps by THIS function.
        A_r(p,d) = w0*x(p,d) + w1*A_r(p-r,d) + w2*A_r(p-r,d-1) + w3*A_r(p-r,d+1) + w4*max_i A_r(p-r,i)
    Out-of-image predecessors and d-1 / d+1 outside [0,D) read 0 (our choice, mirroring the
    SPN op's zero reads, gaterecurrent2dnoind_kernel.cu:67-75); at the first pixel of a scan
    only the w0 term remains.  Output = max over the four directions."""
    B, C, D, H, W = x.shape
    g = guidance.view(B, 4, 5, C, H, W)
    g = g / g.abs().sum(dim=2, keepdim=True).clamp_min(1e-12)      # F.normalize(p=1, dim=tap)
    best = None
    for r in range(4):
        w = g[:, r]                                               # [B,5,C,H,W]
        A = torch.zeros_like(x)
        n = W if r < 2 else H
        order = range(n) if r in (0, 2) else range(n - 1, -1, -1)
        prev = None
        for t in order:
            if r < 2:
                xs = x[..., t]; ws = w[..., t]                     # [B,C,D,H], [B,5,C,H]
                wk = [ws[:, k].unsqueeze(2) for k in range(5)]     # [B,C,1,H]
            else:
                xs = x[:, :, :, t, :]; ws = w[:, :, :, t, :]       # [B,C,D,W], [B,5,C,W]
                wk = [ws[:, k].unsqueeze(2) for k in range(5)]
            cur = wk[0] * xs
            if prev is not None:
                dm = F.pad(prev, (0, 0, 1, 0))[:, :, :D]           # A(p-r, d-1)
                dp = F.pad(prev, (0, 0, 0, 1))[:, :, 1:]           # A(p-r, d+1)
                mx = prev.max(dim=2, keepdim=True).values
                cur = cur + wk[1] * prev + wk[2] * dm + wk[3] * dp + wk[4] * mx
            if r < 2:
                A[..., t] = cur
            else:
                A[:, :, :, t, :] = cur
            prev = cur
        best = A if best is None else torch.maximum(best, A)
    return best


def lga(x, guidance, radius=2):
    """PARITY UNPINNED.  Local guided aggregation layer of GANet (eq. 6).  x: [B,D,H,W]
    (single-channel cost volume); guidance: [B, 3*(2r+1)^2, H, W] ordered (tap-plane t in
    {0: d, 1: d-1, 2: d+1}, ky, kx), L1-normalised over all 75 weights by THIS function:
        A(p,d) = sum_{q in window(p)} w0(p,q) x(q,d) + w1(p,q) x(q,d-1) + w2(p,q) x(q,d+1)
    Out-of-image q and out-of-range d read 0."""
    B, D, H, W = x.shape
    K = 2 * radius + 1
    g = guidance / guidance.abs().sum(dim=1, keepdim=True).clamp_min(1e-12)
    g = g.view(B, 3, K, K, H, W)
    planes = [x, F.pad(x, (0, 0, 0, 0, 1, 0))[:, :D], F.pad(x, (0, 0, 0, 0, 0, 1))[:, 1:]]
    out = torch.zeros_like(x)
    for t in range(3):
        xp = F.pad(planes[t], (radius, radius, radius, radius))
        for ky in range(K):
            for kx in range(K):
                out += g[:, t, ky, kx].unsqueeze(1) * xp[:, :, ky:ky + H, kx:kx + W]
    return out


# --------------------------------------------------------------------------------------
# metric
# --------------------------------------------------------------------------------------
def epe(est, gt, lb=0, ub=192):
    """calc_error(...)['epe'] (data/datasets/evaluation/stereo/pixel_error.py:40-65): mean
    |gt-est| over lb < gt < ub; 0 when the mask is empty."""
    mask = (gt > lb) & (gt < ub)
    if mask.sum() < 1:
        return 0.0
    return float((gt[mask] - est[mask]).abs().mean())


# --------------------------------------------------------------------------------------
# whole hot path (used by bench.py's cpu_baseline / --impl reference)
# --------------------------------------------------------------------------------------
def psm_hot_path(sd, left_fm, right_fm, max_disp=192, prefix="cost_processor.aggregator.", dtype=torch.float32):
    """CatCostProcessor.forward + FasterSoftArgmin per cost (cost_processors/builder.py:33-40,
    models/general_stereo_model.py:51-54) for the PSMNet scene_flow config.  `dtype=torch.float64`
    evaluates the same arithmetic in double precision: the "truth" that two float32 implementations
    (the reference's and ours) are both compared with at sizes where float32 rounding noise in the
    disparity exceeds the 1e-3 px tolerance."""
    raw = cat_volume(left_fm, right_fm, max_disp // 4, 0, 1)
    if dtype != torch.float32:
        raw = raw.to(dtype)
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    costs = psm_aggregator(sd, raw, max_disp, prefix)
    disps = [soft_argmin(c, max_disp) for c in costs]
    return costs, disps


# --------------------------------------------------------------------------------------
# training step of the hot path (BASELINE config 5)
# --------------------------------------------------------------------------------------
def disp_smooth_l1(disps, gt, max_disp, start_disp=0, weights=(1.0, 0.7, 0.5)):
    """DispSmoothL1Loss at full resolution (losses/smooth_l1_loss.py:39-91): per level the mean
    smooth-L1 over start_disp < gt < max_disp, weighted (configs/PSMNet/scene_flow.py:55-63)."""
    mask = (gt > start_disp) & (gt < max_disp)
    total = 0.0
    for w, d in zip(weights, disps):
        total = total + w * F.smooth_l1_loss(d[mask], gt[mask], reduction="mean")
    return total


def train_step(sd, left_fm, right_fm, gt, max_disp, kind="PSMNet", prefix="", weights=(1.0, 0.7, 0.5), shards=1):
    """One forward + backward of cat volume -> aggregator (training-mode BatchNorm) -> soft-argmin ->
    smooth-L1, through torch autograd on the functional restatement above.  Returns the loss, the three
    disparity maps, d(loss)/d(every floating-point parameter), d(loss)/d(features) and the running
    statistics the BatchNorm layers hold afterwards.  `shards` > 1 restates the data-parallel step with
    synchronised BatchNorm: the batch statistics span the whole batch, the loss is the mean over `shards`
    equal batch slices of each slice's own masked mean (what averaging per-rank gradients computes)."""
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    leaf = dict(sd)
    for k in names:
        leaf[k] = sd[k].clone().requires_grad_(True)
    l = left_fm.clone().requires_grad_(True)
    r = right_fm.clone().requires_grad_(True)
    train = {}
    raw = cat_volume(l, r, max_disp // 4, 0, 1)
    agg = acf_aggregator if kind == "AcfNet" else psm_aggregator
    costs = agg(leaf, raw, max_disp, prefix, True, train)
    disps = [soft_argmin(c, max_disp) for c in costs]
    n = gt.shape[0] // shards
    loss = sum(disp_smooth_l1([d[i * n:(i + 1) * n] for d in disps], gt[i * n:(i + 1) * n], max_disp, 0, weights)
               for i in range(shards)) / shards
    loss.backward()
    grads = {k: (leaf[k].grad if leaf[k].grad is not None else torch.zeros_like(sd[k])) for k in names}
    return dict(loss=loss.detach(), disps=[d.detach() for d in disps], grads=grads, dleft=l.grad, dright=r.grad,
                running=train["running"])
