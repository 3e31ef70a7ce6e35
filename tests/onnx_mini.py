"""Test infrastructure: a minimal reader + evaluator for the ONNX files written by wekws_b200.export_onnx.

`onnx` / `onnxruntime` are not in this image, so the exporter's output is checked by decoding the protobuf wire format
again (generic decoder + the field numbers of onnx.proto) and evaluating the opset-13 operators it uses with torch CPU
ops -- the same call keyword_spotting.cc makes through ORT: run(input, cache) -> (output, r_cache)."""
import numpy as np
import torch
import torch.nn.functional as F


def _read_varint(buf, pos):
    v, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return v, pos


def fields(buf):
    """-> list of (field number, wire type, value): varint -> int, 64/32-bit -> bytes, length-delimited -> bytes."""
    out, pos = [], 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v, pos = bytes(buf[pos:pos + n]), pos + n
        else:
            raise ValueError("wire type %d" % wt)
        out.append((fno, wt, v))
    assert pos == len(buf)
    return out


def _signed(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _tensor(buf):
    dims, dt, name, raw = [], None, None, None
    for f, _, v in fields(buf):
        if f == 1:
            dims.append(_signed(v))
        elif f == 2:
            dt = v
        elif f == 8:
            name = v.decode()
        elif f == 9:
            raw = v
    np_dt = {1: "<f4", 7: "<i8"}[dt]
    return name, torch.from_numpy(np.frombuffer(raw, dtype=np_dt).reshape(dims).copy())


def _value_info(buf):
    name, shape = None, []
    for f, _, v in fields(buf):
        if f == 1:
            name = v.decode()
        elif f == 2:
            (tt,) = [x for ff, _, x in fields(v) if ff == 1]
            for ff, _, x in fields(tt):
                if ff == 2:
                    for _, _, dim in fields(x):
                        d = fields(dim)[0]
                        shape.append(d[2].decode() if d[0] == 2 else d[2])
    return name, shape


def load(path):
    """-> dict(ir_version, opset, producer, metadata, inputs, outputs, initializers, nodes)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    m = dict(metadata={}, inputs=[], outputs=[], initializers={}, nodes=[])
    for f, _, v in fields(buf):
        if f == 1:
            m["ir_version"] = v
        elif f == 2:
            m["producer"] = v.decode()
        elif f == 8:
            d = {ff: x for ff, _, x in fields(v)}
            assert d.get(1, b"") == b""
            m["opset"] = d[2]
        elif f == 14:
            d = {ff: x.decode() for ff, _, x in fields(v)}
            m["metadata"][d[1]] = d[2]
        elif f == 7:
            for gf, _, gv in fields(v):
                if gf == 1:
                    node = dict(inputs=[], outputs=[], attrs={})
                    for nf, _, nv in fields(gv):
                        if nf == 1:
                            node["inputs"].append(nv.decode())
                        elif nf == 2:
                            node["outputs"].append(nv.decode())
                        elif nf == 4:
                            node["op"] = nv.decode()
                        elif nf == 5:
                            a = fields(nv)
                            an = [x for ff, _, x in a if ff == 1][0].decode()
                            at = [x for ff, _, x in a if ff == 20][0]
                            if at == 2:
                                node["attrs"][an] = _signed([x for ff, _, x in a if ff == 3][0])
                            elif at == 7:
                                node["attrs"][an] = [_signed(x) for ff, _, x in a if ff == 8]
                            else:
                                raise ValueError("attribute type %d" % at)
                    m["nodes"].append(node)
                elif gf == 5:
                    n, t = _tensor(gv)
                    m["initializers"][n] = t
                elif gf == 11:
                    m["inputs"].append(_value_info(gv))
                elif gf == 12:
                    m["outputs"].append(_value_info(gv))
    return m


@torch.no_grad()
def run(m, feeds):
    """Evaluates the graph in node order (the exporter writes it topologically sorted)."""
    env = dict(m["initializers"])
    env.update(feeds)
    for n in m["nodes"]:
        x = [env[i] for i in n["inputs"]]
        a, op = n["attrs"], n["op"]
        if op == "Sub":
            y = x[0] - x[1]
        elif op == "Mul":
            y = x[0] * x[1]
        elif op == "Add":
            y = x[0] + x[1]
        elif op == "MatMul":
            y = torch.matmul(x[0], x[1])
        elif op == "Relu":
            y = F.relu(x[0])
        elif op == "Sigmoid":
            y = torch.sigmoid(x[0])
        elif op == "Softmax":
            y = torch.softmax(x[0], a["axis"])
        elif op == "Identity":
            y = x[0]
        elif op == "Transpose":
            y = x[0].permute(a["perm"])
        elif op == "Concat":
            y = torch.cat(x, a["axis"])
        elif op == "Squeeze":
            y = x[0].squeeze(int(x[1][0]))
        elif op == "Unsqueeze":
            y = x[0].unsqueeze(int(x[1][0]))
        elif op == "Slice":
            st, en, ax = int(x[1][0]), int(x[2][0]), int(x[3][0])
            size = x[0].size(ax)
            if st < 0:
                st = max(0, st + size)
            en = min(en, size) if en >= 0 else en + size
            y = x[0].narrow(ax, st, max(0, en - st))
        elif op == "Conv":
            assert a["pads"] == [0, 0] and a["strides"] == [1] and a["kernel_shape"] == [x[1].size(2)]
            y = F.conv1d(x[0], x[1], x[2], dilation=a["dilations"][0], groups=a["group"])
        else:
            raise NotImplementedError(op)
        env[n["outputs"][0]] = y
    return tuple(env[name] for name, _ in m["outputs"])
