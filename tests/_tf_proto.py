"""Test helper: TensorFlow's GraphDef schema (the public field numbers of graph.proto, node_def.proto,
attr_value.proto, tensor.proto, tensor_shape.proto, types.proto) declared at run time with google.protobuf, so
that frozen graphs can be SERIALISED by the real protobuf library and fed to keyword_spotting_b200.graph_pb's
independent wire-format reader.  Only the fields the reference's graphs use are declared."""
import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=_F.LABEL_OPTIONAL, type_name=None, packed=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    if packed is not None:
        f.options.packed = packed
    return f


def _build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "kws_test_tf_graph.proto"
    fd.package = "kwstf"
    fd.syntax = "proto3"
    dim = fd.message_type.add(); dim.name = "Dim"
    _field(dim, "size", 1, _F.TYPE_INT64); _field(dim, "name", 2, _F.TYPE_STRING)
    shape = fd.message_type.add(); shape.name = "TensorShapeProto"
    _field(shape, "dim", 2, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".kwstf.Dim"); _field(shape, "unknown_rank", 3, _F.TYPE_BOOL)
    tensor = fd.message_type.add(); tensor.name = "TensorProto"
    _field(tensor, "dtype", 1, _F.TYPE_INT32)
    _field(tensor, "tensor_shape", 2, _F.TYPE_MESSAGE, type_name=".kwstf.TensorShapeProto")
    _field(tensor, "version_number", 3, _F.TYPE_INT32)
    _field(tensor, "tensor_content", 4, _F.TYPE_BYTES)
    _field(tensor, "float_val", 5, _F.TYPE_FLOAT, _F.LABEL_REPEATED, packed=True)
    _field(tensor, "double_val", 6, _F.TYPE_DOUBLE, _F.LABEL_REPEATED, packed=True)
    _field(tensor, "int_val", 7, _F.TYPE_INT32, _F.LABEL_REPEATED, packed=True)
    _field(tensor, "int64_val", 10, _F.TYPE_INT64, _F.LABEL_REPEATED, packed=True)
    attr = fd.message_type.add(); attr.name = "AttrValue"
    _field(attr, "s", 2, _F.TYPE_BYTES); _field(attr, "i", 3, _F.TYPE_INT64); _field(attr, "f", 4, _F.TYPE_FLOAT)
    _field(attr, "b", 5, _F.TYPE_BOOL); _field(attr, "type", 6, _F.TYPE_INT32)
    _field(attr, "shape", 7, _F.TYPE_MESSAGE, type_name=".kwstf.TensorShapeProto")
    _field(attr, "tensor", 8, _F.TYPE_MESSAGE, type_name=".kwstf.TensorProto")
    entry = fd.message_type.add(); entry.name = "AttrEntry"          # wire-identical to map<string, AttrValue>
    _field(entry, "key", 1, _F.TYPE_STRING); _field(entry, "value", 2, _F.TYPE_MESSAGE, type_name=".kwstf.AttrValue")
    node = fd.message_type.add(); node.name = "NodeDef"
    _field(node, "name", 1, _F.TYPE_STRING); _field(node, "op", 2, _F.TYPE_STRING)
    _field(node, "input", 3, _F.TYPE_STRING, _F.LABEL_REPEATED); _field(node, "device", 4, _F.TYPE_STRING)
    _field(node, "attr", 5, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".kwstf.AttrEntry")
    ver = fd.message_type.add(); ver.name = "VersionDef"
    _field(ver, "producer", 1, _F.TYPE_INT32)
    graph = fd.message_type.add(); graph.name = "GraphDef"
    _field(graph, "node", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".kwstf.NodeDef")
    _field(graph, "versions", 4, _F.TYPE_MESSAGE, type_name=".kwstf.VersionDef")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    if get is None:                                                   # older protobuf
        factory = message_factory.MessageFactory(pool)
        get = factory.GetPrototype
    return {n: get(pool.FindMessageTypeByName("kwstf." + n)) for n in
            ("GraphDef", "NodeDef", "AttrValue", "TensorProto", "TensorShapeProto")}


M = _build()
DT = {np.dtype(np.float32): 1, np.dtype(np.int32): 3, np.dtype(np.int8): 6, np.dtype(np.float64): 2, np.dtype(np.int64): 9}


def tensor_proto(arr, dtype_enum=None, as_content=True):
    arr = np.asarray(arr)
    t = M["TensorProto"]()
    t.dtype = dtype_enum if dtype_enum is not None else DT[arr.dtype]
    for d in arr.shape:
        t.tensor_shape.dim.add().size = int(d)
    if as_content:
        t.tensor_content = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
    elif arr.dtype == np.float32:
        t.float_val.extend(float(x) for x in arr.ravel())
    elif arr.dtype == np.float64:
        t.double_val.extend(float(x) for x in arr.ravel())
    elif arr.dtype == np.int64:
        t.int64_val.extend(int(x) for x in arr.ravel())
    else:
        t.int_val.extend(int(x) for x in arr.ravel())
    return t


def add_node(graph, name, op, inputs=(), **attrs):
    n = graph.node.add()
    n.name, n.op = name, op
    n.input.extend(inputs)
    for k, v in attrs.items():
        e = n.attr.add()
        e.key = k
        if isinstance(v, bool):
            e.value.b = v
        elif isinstance(v, float):
            e.value.f = v
        elif isinstance(v, int):
            e.value.i = v
        elif isinstance(v, bytes):
            e.value.s = v
        elif isinstance(v, tuple) and v and v[0] == "type":
            e.value.type = v[1]
        else:
            e.value.tensor.CopyFrom(v)
    return n
