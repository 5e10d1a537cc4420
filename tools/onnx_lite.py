"""Minimal ONNX protobuf writer/reader (no `onnx` package in this image).

Only the subset of onnx.proto3 that the april-asr exporter's three graphs use
(SURVEY.md Appendix B): ModelProto{ir_version=1, opset_import=8, graph=7},
GraphProto{node=1,name=2,initializer=5,input=11,output=12},
NodeProto{input=1,output=2,name=3,op_type=4,attribute=5},
AttributeProto{name=1,f=2,i=3,s=4,t=5,floats=7,ints=8,type=20},
TensorProto{dims=1,data_type=2,name=8,raw_data=9},
ValueInfoProto{name=1,type=2{tensor_type=1{elem_type=1,shape=2{dim=1{dim_value=1}}}}}.

Test/fixture infrastructure only; the product's reader is C++ (csrc/onnx_reader.cpp).
"""
import struct

import numpy as np

FLOAT, INT32, INT64 = 1, 6, 7
_NP2ONNX = {np.dtype("float32"): FLOAT, np.dtype("int64"): INT64, np.dtype("int32"): INT32}
_ONNX2NP = {FLOAT: np.float32, INT64: np.int64, INT32: np.int32}

ATTR_FLOAT, ATTR_INT, ATTR_STRING, ATTR_TENSOR, ATTR_FLOATS, ATTR_INTS = 1, 2, 3, 4, 6, 7


# ----------------------------------------------------------------------------- wire helpers
def _varint(v):
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field, wt):
    return _varint((field << 3) | wt)


def f_varint(field, v):
    return _key(field, 0) + _varint(int(v))


def f_bytes(field, b):
    if isinstance(b, str):
        b = b.encode("utf-8")
    return _key(field, 2) + _varint(len(b)) + bytes(b)


def f_float(field, v):
    return _key(field, 5) + struct.pack("<f", v)


# ----------------------------------------------------------------------------- writer
def tensor_proto(name, arr):
    arr = np.ascontiguousarray(arr)
    out = b""
    for d in arr.shape:
        out += f_varint(1, d)
    out += f_varint(2, _NP2ONNX[arr.dtype])
    out += f_bytes(8, name)
    out += f_bytes(9, arr.tobytes())
    return out


def attribute(name, value):
    out = f_bytes(1, name)
    if isinstance(value, float):
        out += f_float(2, value) + f_varint(20, ATTR_FLOAT)
    elif isinstance(value, (int, np.integer)):
        out += f_varint(3, value) + f_varint(20, ATTR_INT)
    elif isinstance(value, (bytes, str)):
        out += f_bytes(4, value) + f_varint(20, ATTR_STRING)
    elif isinstance(value, np.ndarray):
        out += f_bytes(5, tensor_proto("", value)) + f_varint(20, ATTR_TENSOR)
    elif isinstance(value, (list, tuple)):
        if len(value) and isinstance(value[0], float):
            for v in value:
                out += f_float(7, v)
            out += f_varint(20, ATTR_FLOATS)
        else:
            for v in value:
                out += f_varint(8, v)
            out += f_varint(20, ATTR_INTS)
    else:
        raise TypeError(type(value))
    return out


def node(op_type, inputs, outputs, name="", **attrs):
    out = b""
    for i in inputs:
        out += f_bytes(1, i)
    for o in outputs:
        out += f_bytes(2, o)
    out += f_bytes(3, name or (op_type + "_" + outputs[0]))
    out += f_bytes(4, op_type)
    for k, v in attrs.items():
        out += f_bytes(5, attribute(k, v))
    return out


def value_info(name, elem_type, shape):
    dims = b""
    for d in shape:
        dims += f_bytes(1, f_varint(1, d))
    tensor_type = f_varint(1, elem_type) + f_bytes(2, dims)
    return f_bytes(1, name) + f_bytes(2, f_bytes(1, tensor_type))


def model(graph_name, nodes, initializers, inputs, outputs, opset=11):
    parts = [f_bytes(1, n) for n in nodes]                # joined once: repeated += on a 300 MB bytes object is quadratic
    parts.append(f_bytes(2, graph_name))
    parts += [f_bytes(5, t) for t in initializers]
    parts += [f_bytes(11, vi) for vi in inputs]
    parts += [f_bytes(12, vi) for vi in outputs]
    g = b"".join(parts)
    m = f_varint(1, 6)                       # ir_version 6 (what torch 1.13 emits for opset 11)
    m += f_bytes(2, "pytorch") + f_bytes(3, "1.13.0")
    m += f_bytes(7, g)
    m += f_bytes(8, f_bytes(1, "") + f_varint(2, opset))
    return m


class GraphBuilder:
    """Tiny helper that hands out fresh value names the way torch.onnx does."""

    def __init__(self, name):
        self.name = name
        self.nodes, self.inits, self.inputs, self.outputs = [], [], [], []
        self._n = 0

    def fresh(self, prefix="t"):
        self._n += 1
        return "%s_%d" % (prefix, self._n)

    def init(self, arr, name=None):
        name = name or self.fresh("onnx::Const")
        self.inits.append(tensor_proto(name, np.asarray(arr)))
        return name

    def const(self, arr):
        """A Constant node (how torch exports scalars)."""
        out = self.fresh("const")
        self.nodes.append(node("Constant", [], [out], value=np.asarray(arr)))
        return out

    def op(self, op_type, inputs, n_out=1, out_names=None, **attrs):
        outs = out_names or [self.fresh(op_type.lower()) for _ in range(n_out)]
        self.nodes.append(node(op_type, inputs, outs, **attrs))
        return outs[0] if len(outs) == 1 else outs

    def add_input(self, name, elem_type, shape):
        self.inputs.append(value_info(name, elem_type, shape))

    def add_output(self, name, elem_type, shape):
        self.outputs.append(value_info(name, elem_type, shape))

    def serialize(self):
        return model(self.name, self.nodes, self.inits, self.inputs, self.outputs)


# ----------------------------------------------------------------------------- reader
def _read_varint(buf, pos):
    shift = 0
    v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not (b & 0x80):
            return v, pos
        shift += 7


def parse_message(buf):
    """-> list of (field, wiretype, value) with value int | bytes."""
    pos, out = 0, []
    n = len(buf)
    while pos < n:
        key, pos = _read_varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("wire type %d" % wt)
        out.append((field, wt, v))
    return out


def _s64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_tensor(buf):
    dims, dtype, name, raw = [], FLOAT, "", b""
    floats, int64s = [], []
    for f, wt, v in parse_message(buf):
        if f == 1:
            if wt == 0:
                dims.append(_s64(v))
            else:                            # packed
                p = 0
                while p < len(v):
                    d, p = _read_varint(v, p)
                    dims.append(_s64(d))
        elif f == 2:
            dtype = v
        elif f == 8:
            name = bytes(v).decode()
        elif f == 9:
            raw = bytes(v)
        elif f == 4:
            floats.append(v)
        elif f == 7:
            int64s.append(v)
    if raw or not (floats or int64s):
        arr = np.frombuffer(raw, dtype=_ONNX2NP[dtype]).reshape(dims).copy()
    elif dtype == FLOAT:
        arr = np.frombuffer(b"".join(bytes(x) for x in floats), dtype=np.float32).reshape(dims).copy()
    else:
        arr = np.array([_s64(x) for x in int64s], dtype=np.int64).reshape(dims)
    return name, arr


def parse_model(buf):
    """-> dict(nodes=[{op,inputs,outputs,attrs}], inits={name: ndarray}, inputs=[(name,shape)], outputs=[...])"""
    buf = bytes(buf)
    graph = None
    for f, wt, v in parse_message(buf):
        if f == 7:
            graph = v
    res = dict(nodes=[], inits={}, inputs=[], outputs=[])

    def vinfo(b):
        name, shape, et = "", [], 0
        for f, wt, v in parse_message(b):
            if f == 1:
                name = bytes(v).decode()
            elif f == 2:
                for f2, _, v2 in parse_message(v):
                    if f2 == 1:
                        for f3, _, v3 in parse_message(v2):
                            if f3 == 1:
                                et = v3
                            elif f3 == 2:
                                for f4, _, v4 in parse_message(v3):
                                    if f4 == 1:
                                        for f5, _, v5 in parse_message(v4):
                                            if f5 == 1:
                                                shape.append(_s64(v5))
        return name, et, shape

    for f, wt, v in parse_message(graph):
        if f == 1:
            nd = dict(op="", inputs=[], outputs=[], attrs={}, name="")
            for f2, wt2, v2 in parse_message(v):
                if f2 == 1:
                    nd["inputs"].append(bytes(v2).decode())
                elif f2 == 2:
                    nd["outputs"].append(bytes(v2).decode())
                elif f2 == 3:
                    nd["name"] = bytes(v2).decode()
                elif f2 == 4:
                    nd["op"] = bytes(v2).decode()
                elif f2 == 5:
                    an, val, ints, floats = "", None, [], []
                    for f3, wt3, v3 in parse_message(v2):
                        if f3 == 1:
                            an = bytes(v3).decode()
                        elif f3 == 2:
                            val = struct.unpack("<f", bytes(v3))[0]
                        elif f3 == 3:
                            val = _s64(v3)
                        elif f3 == 4:
                            val = bytes(v3)
                        elif f3 == 5:
                            val = parse_tensor(v3)[1]
                        elif f3 == 7:
                            floats.append(struct.unpack("<f", bytes(v3))[0])
                        elif f3 == 8:
                            if wt3 == 0:
                                ints.append(_s64(v3))
                            else:
                                p = 0
                                while p < len(v3):
                                    d, p = _read_varint(v3, p)
                                    ints.append(_s64(d))
                    if val is None:
                        val = ints if ints else floats
                    nd["attrs"][an] = val
            res["nodes"].append(nd)
        elif f == 5:
            name, arr = parse_tensor(v)
            res["inits"][name] = arr
        elif f == 11:
            res["inputs"].append(vinfo(v))
        elif f == 12:
            res["outputs"].append(vinfo(v))
    return res
