"""Weight formats: frozen graph.pb / graph_octbit.pb readers against GraphDefs serialised by
the real protobuf library from TensorFlow's public schema (tests/_tf_proto.py)."""
import numpy as np
import pytest

from oracle import model as om, octbit as ooct
from tests import _tf_proto as tp


def _frozen_graph(ow, naming="kernel", octbit=False, octbit_fc=False):
    """A frozen rnn_ctc deployment graph the way main.py:339-348 writes it (Consts + read Identities + ops)."""
    g = tp.M["GraphDef"]()
    g.versions.producer = 22
    kname, bname = ("kernel", "bias") if naming == "kernel" else ("weights", "biases")
    tp.add_node(g, "model/inputX", "Placeholder", dtype=("type", 1))
    tp.add_node(g, "model/rnn_initial_states", "Placeholder", dtype=("type", 1))
    tp.add_node(g, "model/Const", "Const", value=tp.tensor_proto(ow.mel_basis), dtype=("type", 1))
    tp.add_node(g, "model/ExpandDims_1/dim", "Const", value=tp.tensor_proto(np.asarray(0, np.int32), as_content=False), dtype=("type", 3))
    tp.add_node(g, "model/frame/range/delta", "Const", value=tp.tensor_proto(np.asarray([160], np.int32), as_content=False), dtype=("type", 3))
    octs = []
    for l in range(ow.num_layers):
        base = "model/drnn/multi_rnn_cell/cell_%d/gru_cell" % l
        for part, kern, bias in (("gates", ow.gates_kernel[l], ow.gates_bias[l]), ("candidate", ow.cand_kernel[l], ow.cand_bias[l])):
            # biases as float_val lists, a constant-filled one as the single-value short form
            b_proto = tp.tensor_proto(bias, as_content=False)
            if np.all(bias == bias[0]):
                b_proto = tp.tensor_proto(bias[:1], as_content=False)
                del b_proto.tensor_shape.dim[:]
                b_proto.tensor_shape.dim.add().size = len(bias)
            tp.add_node(g, "%s/%s/%s" % (base, part, bname), "Const", value=b_proto, dtype=("type", 1))
            tp.add_node(g, "%s/%s/%s/read" % (base, part, bname), "Identity", ["%s/%s/%s" % (base, part, bname)])
            wname = "%s/%s/%s" % (base, part, kname)
            mm = "model/drnn/while/multi_rnn_cell/cell_%d/gru_cell/%s/MatMul" % (l, part)
            if octbit and l > 0:                       # default_octbit_matmul_name_check skips cell_0
                wq, scale, qb = ooct.octize_weight_int8_signed(kern)
                tp.add_node(g, wname, "Const", value=tp.tensor_proto(wq, dtype_enum=11), dtype=("type", 11))
                tp.add_node(g, wname + "/read", "Identity", [wname])
                tp.add_node(g, mm + "/Enter", "Enter", [wname + "/read"])
                tp.add_node(g, mm, "OctbitMatMul", ["model/drnn/while/concat_%d" % l, mm + "/Enter"],
                            transpose_a=False, transpose_b=True, scale=float(np.float32(scale)),
                            bias=tp.tensor_proto(qb.astype(np.float32)))
                octs.append((mm, wq, np.float32(scale), qb.astype(np.float32)))
            else:
                tp.add_node(g, wname, "Const", value=tp.tensor_proto(kern), dtype=("type", 1))
                tp.add_node(g, wname + "/read", "Identity", [wname])
                tp.add_node(g, mm, "MatMul", ["model/drnn/while/concat_%d" % l, wname + "/read"])
    if octbit and octbit_fc:                           # "model/MatMul_1" passes default_octbit_matmul_name_check too
        wq, scale, qb = ooct.octize_weight_int8_signed(ow.fc_w)
        tp.add_node(g, "model/weightsClasses", "Const", value=tp.tensor_proto(wq, dtype_enum=11), dtype=("type", 11))
        tp.add_node(g, "model/weightsClasses/read", "Identity", ["model/weightsClasses"])
        tp.add_node(g, "model/MatMul_1", "OctbitMatMul", ["model/Reshape", "model/weightsClasses/read"],
                    transpose_a=False, transpose_b=True, scale=float(np.float32(scale)), bias=tp.tensor_proto(qb.astype(np.float32)))
        octs.append(("model/MatMul_1", wq, np.float32(scale), qb.astype(np.float32)))
    else:
        tp.add_node(g, "model/weightsClasses", "Const", value=tp.tensor_proto(ow.fc_w), dtype=("type", 1))
    tp.add_node(g, "model/biasesClasses", "Const", value=tp.tensor_proto(ow.fc_b, as_content=False), dtype=("type", 1))
    tp.add_node(g, "model/softmax", "Softmax", ["model/test"])
    return g.SerializeToString(), octs


@pytest.mark.parametrize("naming", ["kernel", "weights"])
@pytest.mark.parametrize("n_mel", [40, 60])
def test_frozen_graph_weights_round_trip(naming, n_mel):
    from keyword_spotting_b200 import graph_pb
    ow = om.init_weights(seed=11, n_mel=n_mel)
    data, _ = _frozen_graph(ow, naming)
    nodes = graph_pb.load_graph(data)
    assert {n.name for n in nodes} >= {"model/inputX", "model/rnn_initial_states", "model/softmax"}
    w = graph_pb.rnn_ctc_weights(nodes, n_mel=n_mel)
    np.testing.assert_array_equal(w.mel_basis, ow.mel_basis)
    for l in range(2):
        np.testing.assert_array_equal(w.gates_kernel[l], ow.gates_kernel[l])
        np.testing.assert_array_equal(w.gates_bias[l], ow.gates_bias[l])
        np.testing.assert_array_equal(w.cand_kernel[l], ow.cand_kernel[l])
        np.testing.assert_array_equal(w.cand_bias[l], ow.cand_bias[l])
    np.testing.assert_array_equal(w.fc_w, ow.fc_w)
    np.testing.assert_array_equal(w.fc_b, ow.fc_b)
    consts = graph_pb.constants(nodes)
    assert consts["model/ExpandDims_1/dim"].shape == () and int(consts["model/ExpandDims_1/dim"]) == 0
    assert consts["model/frame/range/delta"].tolist() == [160]


def test_octbit_graph_nodes():
    from keyword_spotting_b200 import graph_pb
    ow = om.init_weights(seed=12, n_mel=40)
    data, octs = _frozen_graph(ow, "kernel", octbit=True)
    got = graph_pb.octbit_nodes(graph_pb.load_graph(data))
    assert [o.name for o in got] == [o[0] for o in octs] and len(got) == 2
    for o, (_, wq, scale, qb) in zip(got, octs):
        assert o.weight_q.dtype == np.int8
        np.testing.assert_array_equal(o.weight_q, wq)
        assert np.float32(o.scale) == scale
        np.testing.assert_array_equal(o.bias, qb)
    # layer 0 stays a float MatMul, so the float weights of layer 0 are still there but layer 1's are not
    with pytest.raises(graph_pb.GraphFormatError):
        graph_pb.rnn_ctc_weights(graph_pb.load_graph(data), n_mel=40)


@pytest.mark.parametrize("octbit_fc", [True, False])
def test_octbit_graph_to_model_weights(octbit_fc):
    """graph_octbit.pb -> (ModelWeights, OctbitModelWeights): cell_0 and the biases stay float and exact, converted
    MatMuls carry the rewriter's qint8 / scale / bias, their float image is the dequantised kernel."""
    from keyword_spotting_b200 import graph_pb
    ow = om.init_weights(seed=12, n_mel=40)
    data, octs = _frozen_graph(ow, "kernel", octbit=True, octbit_fc=octbit_fc)
    w, octw = graph_pb.rnn_ctc_octbit_weights(graph_pb.load_graph(data), n_mel=40)
    np.testing.assert_array_equal(w.gates_kernel[0], ow.gates_kernel[0])
    np.testing.assert_array_equal(w.cand_kernel[0], ow.cand_kernel[0])
    for l in range(2):
        np.testing.assert_array_equal(w.gates_bias[l], ow.gates_bias[l])
        np.testing.assert_array_equal(w.cand_bias[l], ow.cand_bias[l])
    np.testing.assert_array_equal(w.fc_b, ow.fc_b)
    assert sorted(octw.gates) == sorted(octw.candidate) == [1]
    want = om.octize_model(ow, fc=octbit_fc)
    for got, exp in ((octw.gates[1], want["gates"][1]), (octw.candidate[1], want["candidate"][1])):
        np.testing.assert_array_equal(got.weight_q, exp[0])
        assert np.float32(got.scale) == np.float32(exp[1])
        np.testing.assert_array_equal(got.bias, exp[2])
    assert (octw.fc is not None) == octbit_fc
    if octbit_fc:
        np.testing.assert_array_equal(octw.fc.weight_q, want["fc"][0])
        assert np.abs(w.fc_w - ow.fc_w).max() <= np.float32(want["fc"][1]) * 0.5 + 1e-7      # dequantised image
    else:
        np.testing.assert_array_equal(w.fc_w, ow.fc_w)
    assert np.abs(w.gates_kernel[1] - ow.gates_kernel[1]).max() <= np.float32(want["gates"][1][1]) * 0.5 + 1e-7
    # a float graph is not an octbit graph
    with pytest.raises(graph_pb.GraphFormatError):
        graph_pb.rnn_ctc_octbit_weights(graph_pb.load_graph(_frozen_graph(ow)[0]), n_mel=40)


def test_malformed_graphs_fail_loudly():
    from keyword_spotting_b200 import graph_pb
    with pytest.raises(graph_pb.GraphFormatError):
        graph_pb.load_graph(b"")
    with pytest.raises(graph_pb.GraphFormatError):
        graph_pb.load_graph(b"\x0a\xff\xff\xff\x0f")            # length runs past the buffer
    g = tp.M["GraphDef"]()
    tp.add_node(g, "model/softmax", "Softmax")
    with pytest.raises(graph_pb.GraphFormatError):
        graph_pb.rnn_ctc_weights(graph_pb.load_graph(g.SerializeToString()))
