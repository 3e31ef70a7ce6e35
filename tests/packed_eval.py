"""Test helper: interprets the PACKED weight program of a native model (what csrc/model_host.cu
produced for the kernels) with plain torch ops on the CPU, walking the blobs in the order the
fused kernels consume them.  Validates BN folding + packing without a GPU.  Test code only."""
import ctypes as C

import torch
import torch.nn.functional as F


def pad4(n):
    return (n + 3) & ~3


def read_packed(native, handle):
    lib = native.lib()
    out = []
    for which in (0, 1):
        n = lib.wekws_model_packed_floats(handle, which)
        buf = torch.empty(max(n, 1), dtype=torch.float32)
        native.check(lib.wekws_model_packed_copy(handle, which, C.c_void_p(buf.data_ptr()), buf.numel()), "packed_copy")
        out.append(buf[:n])
    return out


def eval_conv(stream, vec, kind, Cc, idim, odim, K, dils, stack_size, act_sigmoid, has_cmvn, x, cache=None):
    B, T, _ = x.shape
    pos = [0]

    def take(rows):
        w = stream[pos[0]:pos[0] + rows * Cc].reshape(rows, Cc)
        pos[0] += rows * Cc
        return w

    v_mean, v_istd = 0, pad4(idim)
    v_bp = 2 * pad4(idim)
    v_blocks = v_bp + Cc
    stride = {"mdtc": (K + 3) * Cc, "ds_tcn": (K + 2) * Cc, "tcn": Cc}[kind]
    v_wc = v_blocks + len(dils) * stride
    v_bc = v_wc + pad4(Cc * odim)
    if has_cmvn:
        x = (x - vec[v_mean:v_mean + idim]) * vec[v_istd:v_istd + idim]
    h = F.relu(x @ take(idim) + vec[v_bp:v_bp + Cc]).transpose(1, 2)          # (B, C, T)
    off, caches, msum = 0, [], torch.zeros(B, Cc, T)
    for bi, d in enumerate(dils):
        pad = d * (K - 1)
        vb = vec[v_blocks + bi * stride: v_blocks + (bi + 1) * stride]
        cin = torch.zeros(B, Cc, pad) if cache is None else cache[:, :, off:off + pad]
        cat = torch.cat((cin, h), dim=2)
        caches.append(cat[:, :, -pad:])
        off += pad
        if kind == "tcn":
            acc = torch.zeros(B, T, Cc)
            for j in range(K):
                acc = acc + cat[:, :, j * d: j * d + T].transpose(1, 2) @ take(Cc)
            h = F.relu(acc + vb[:Cc]).transpose(1, 2) + h
            continue
        dw = vb[:K * Cc].reshape(K, Cc)
        a = vb[K * Cc:(K + 1) * Cc].reshape(1, Cc, 1).expand(B, Cc, T).clone()
        for j in range(K):
            a = a + dw[j].reshape(1, Cc, 1) * cat[:, :, j * d: j * d + T]
        if kind == "ds_tcn":
            a = F.relu(a)
        o = a.transpose(1, 2) @ take(Cc) + vb[(K + 1) * Cc:(K + 2) * Cc]
        if kind == "ds_tcn":
            h = F.relu(o).transpose(1, 2) + h
        else:
            o = F.relu(o) @ take(Cc) + vb[(K + 2) * Cc:(K + 3) * Cc]
            h = F.relu(o.transpose(1, 2) + h)
            if bi > 0 and bi % stack_size == 0:
                msum = msum + h
    assert pos[0] == stream.numel(), "kernel and packer disagree on the weight-stream length"
    feat = (msum if kind == "mdtc" else h).transpose(1, 2)
    y = feat @ vec[v_wc:v_wc + Cc * odim].reshape(Cc, odim) + vec[v_bc:v_bc + odim]
    return (torch.sigmoid(y) if act_sigmoid else y), torch.cat(caches, dim=2)


def eval_gru(vec, H, L, idim, odim, act_sigmoid, has_cmvn, x, h0):
    G = 3 * H
    v_mean, v_istd = 0, pad4(idim)
    v_wp = 2 * pad4(idim)
    v_bp = v_wp + idim * H
    v_layers = v_bp + H
    stride = 2 * H * G + 2 * G
    v_wc = v_layers + L * stride
    v_bc = v_wc + pad4(H * odim)
    if has_cmvn:
        x = (x - vec[v_mean:v_mean + idim]) * vec[v_istd:v_istd + idim]
    xin = F.relu(x @ vec[v_wp:v_wp + idim * H].reshape(idim, H) + vec[v_bp:v_bp + H])
    h = [h0[l].clone() for l in range(L)]
    ys = []
    for t in range(x.shape[1]):
        inp = xin[:, t]
        for l in range(L):
            base = v_layers + l * stride
            w2 = vec[base:base + 2 * H * G].reshape(H, 2 * G)       # row k: [W_ih[:, k] | W_hh[:, k]]
            wih, whh = w2[:, :G], w2[:, G:]
            gi = inp @ wih + vec[base + 2 * H * G: base + 2 * H * G + G]
            gh = h[l] @ whh + vec[base + 2 * H * G + G: base + 2 * H * G + 2 * G]
            r = torch.sigmoid(gi[:, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
            h[l] = (1 - z) * n + z * h[l]
            inp = h[l]
        y = inp @ vec[v_wc:v_wc + H * odim].reshape(H, odim) + vec[v_bc:v_bc + odim]
        ys.append(torch.sigmoid(y) if act_sigmoid else y)
    return torch.stack(ys, dim=1), torch.stack(h, dim=0)
