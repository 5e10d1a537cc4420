"""CPU tests of the .april / ONNX loader (model_loader.cpp; replaces src/file/model_file.c:58-185, src/params.c:46-112 and
ONNXRuntime's model loading for the engine). The real exporter's graphs are not available in this image, so the extraction is
exercised against TWO independently structured writers of the same network, and against crafted malformed graphs."""
import ctypes
import os
import struct

import numpy as np
import pytest

import common as C
import make_synthetic_april as M
import onnx_lite as ox


def _digest(path):
    out = (ctypes.c_uint64 * 16)()
    rc = C.product_lib().c.aab_debug_model_digest(str(path).encode(), out)
    return rc, list(out)


@pytest.mark.parametrize("preset", ["tiny", "mid"])
def test_weight_extraction_is_topology_driven(tmp_path, preset):
    """Style a: MatMul + Add(bias, y) linears, Gemm gate projections (W_ih first), Constant-node scalars, torch-like names.
    Style b: Gemm linears on squeezed operands, MatMul + Add(y, bias) gate projections with W_hh BEFORE W_ih, W_hr as Gemm, all
    initializers anonymous, scalars as initializers, no Identity nodes. Same weights -> the loader must extract the same arrays."""
    a, b = tmp_path / "a.april", tmp_path / "b.april"
    M.generate(str(a), preset, 77, calibrated=False, style="a")
    M.generate(str(b), preset, 77, calibrated=False, style="b")
    assert open(a, "rb").read() != open(b, "rb").read()
    ra, da = _digest(a)
    rb, db = _digest(b)
    assert ra == 0 and rb == 0
    assert da == db
    d = M.PRESETS[preset]
    assert da[:8] == [d["L"], d["D"], d["H"], d["F"], d["J"], d["E"], d["V"], d["CTX"]]
    c, _ = tmp_path / "c.april", M.generate(str(tmp_path / "c.april"), preset, 78, calibrated=False, style="b")
    assert _digest(c)[1][15] != da[15]                      # the hash does depend on the weights


def test_oracle_agrees_on_both_graph_styles(tmp_path):
    """The checker itself (generic ONNX interpreter) must give the same callbacks for both writers."""
    if C.oracle_lib() is None:
        pytest.skip("oracle/_ref not present")
    b = tmp_path / "b.april"
    M.generate(str(b), "tiny", 1234, style="b")
    pcm = C.gen_stream(2, 64000)
    assert C.run_oracle(str(b), pcm).keys() == C.run_oracle(C.TINY_MODEL, pcm).keys()


def _crafted(tmp_path, name, mutate_nodes=None, params_patch=None, mutate_inits=None, inputs_patch=None):
    """tiny model (style a) with a mutation applied to the encoder graph's node list / initializers / inputs, or to the params."""
    d = M.PRESETS["tiny"]
    rs = np.random.RandomState(5)
    vocab = M.make_vocab(d["V"], rs)
    W = M.make_weights(d, 5)
    orig = ox.GraphBuilder.serialize

    def patched(self):
        if self.inputs and len(self.inputs) == 3:                      # the encoder graph
            if mutate_nodes: mutate_nodes(self)
            if mutate_inits: mutate_inits(self)
            if inputs_patch: inputs_patch(self)
        return orig(self)
    ox.GraphBuilder.serialize = patched
    try:
        nets = [M.build_encoder(W, d), M.build_decoder(W, d), M.build_joiner(W, d)]
    finally:
        ox.GraphBuilder.serialize = orig
    pb = M.params_block(d, vocab)
    if params_patch:
        pb = params_patch(pb)
    p = tmp_path / (name + ".april")
    M.write_april(str(p), nets, pb, "crafted", "crafted")
    return p


def _node_index(g, op, k=0):
    """index of the k-th node with op_type `op` in a GraphBuilder (nodes are serialized NodeProtos: field 4 = op_type)."""
    hits = [i for i, n in enumerate(g.nodes) if ox.f_bytes(4, op) in n]
    return hits[k]


def test_crafted_graphs_are_rejected_cleanly(tmp_path):
    """Malformed-but-parseable graphs: nodes with too few inputs, non-positive or absurd dims, wrong dtypes,
    degenerate front-end parameters. Every one must come back as a clean -1 (no crash, no abort, no huge allocation)."""
    cases = {}
    assert _digest(_crafted(tmp_path, "good"))[0] == 0

    def conv_one_input(g):
        i = _node_index(g, "Conv", 1)
        g.nodes[i] = ox.node("Conv", ["x_only"], ["conv_out_x"], strides=[2, 2], kernel_shape=[3, 3])
    cases["conv with one input"] = dict(mutate_nodes=conv_one_input)

    def gemm_one_input(g):
        i = _node_index(g, "Gemm", 0)
        g.nodes[i] = ox.node("Gemm", ["lonely"], ["gemm_out_x"], transB=1)
    cases["gemm with one input"] = dict(mutate_nodes=gemm_one_input)

    def matmul_no_output(g):
        i = _node_index(g, "MatMul", 0)
        g.nodes[i] = ox.f_bytes(1, "a") + ox.f_bytes(1, "b") + ox.f_bytes(4, "MatMul")
    cases["matmul without output"] = dict(mutate_nodes=matmul_no_output)

    def zero_layers(g):
        g.inputs[1] = ox.value_info("h", ox.FLOAT, [0, 1, 64]); g.inputs[2] = ox.value_info("c", ox.FLOAT, [0, 1, 128])
    cases["zero layers"] = dict(inputs_patch=zero_layers)

    def negative_dim(g):
        g.inputs[0] = ox.value_info("x", ox.FLOAT, [1, 9, -80])
    cases["negative mel dim"] = dict(inputs_patch=negative_dim)

    def huge_dim(g):
        g.inputs[1] = ox.value_info("h", ox.FLOAT, [2, 1, 1 << 40])
    cases["huge h dim"] = dict(inputs_patch=huge_dim)

    def int_conv_weight(g):
        for k, t in enumerate(g.inits):
            if ox.f_bytes(8, "encoder.encoder_embed.conv.0.weight") in t:
                g.inits[k] = ox.tensor_proto("encoder.encoder_embed.conv.0.weight", np.zeros((8, 1, 3, 3), np.int64))
    cases["int64 conv weight"] = dict(mutate_inits=int_conv_weight)

    def low_sample_rate(pb):                                            # frame_shift_ms * sample_rate / 1000 == 0
        b = bytearray(pb); b[8 + 4 * 4:8 + 5 * 4] = struct.pack("<i", 50); return bytes(b)
    cases["frame shift below one sample"] = dict(params_patch=low_sample_rate)

    def neg_tensor_dims(g):
        w = np.zeros((4, 4), np.float32)
        t = ox.f_varint(1, (1 << 64) - 4) + ox.f_varint(1, (1 << 64) - 4) + ox.f_varint(2, 1) + ox.f_bytes(8, "weird") + ox.f_bytes(9, w.tobytes())
        g.inits.append(t)
    cases["initializer with negative dims"] = dict(mutate_inits=neg_tensor_dims)

    for name, kw in cases.items():
        p = _crafted(tmp_path, name.replace(" ", "_"), **kw)
        rc, _ = _digest(p)
        assert rc == -1, "crafted model '%s' was accepted" % name
