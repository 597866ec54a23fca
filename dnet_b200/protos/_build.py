"""Build protobuf message classes from hand-written FileDescriptorProtos.

The reference generates ``*_pb2.py`` with grpc_tools (scripts/generate_protos.py:16), which
is not installed here and there is no ``protoc`` binary; the generated files are git-ignored
upstream.  Field numbers, types, labels and proto3 ``optional`` presence below are taken
from src/dnet/protos/dnet_ring.proto:23-113 and shard_api_comm.proto:15-54, so the bytes on
the wire are identical to what a reference peer produces / parses.
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2 as dp
from google.protobuf import descriptor_pool, message_factory

_T = dp.FieldDescriptorProto
_TYPES = {"bytes": _T.TYPE_BYTES, "int32": _T.TYPE_INT32, "int64": _T.TYPE_INT64, "uint64": _T.TYPE_UINT64,
          "string": _T.TYPE_STRING, "bool": _T.TYPE_BOOL, "float": _T.TYPE_FLOAT}


def _msg(fd, name, fields):
    """fields: (name, number, type, flags) with flags in {"", "repeated", "optional"};
    type may be a message type name like ".pkg.Msg"."""
    m = fd.message_type.add()
    m.name = name
    n_opt = 0
    for fname, num, ftype, flag in fields:
        f = m.field.add()
        f.name, f.number = fname, num
        f.label = _T.LABEL_REPEATED if flag == "repeated" else _T.LABEL_OPTIONAL
        if ftype.startswith("."):
            f.type = _T.TYPE_MESSAGE
            f.type_name = ftype
        else:
            f.type = _TYPES[ftype]
        if flag == "optional":          # proto3 explicit presence = synthetic oneof
            f.proto3_optional = True
            od = m.oneof_decl.add()
            od.name = "_" + fname
            f.oneof_index = n_opt
            n_opt += 1
    return m


def build_ring():
    fd = dp.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "dnet_ring.proto", "dnetring", "proto3"
    _msg(fd, "Activation", [("data", 1, "bytes", ""), ("batch_size", 2, "int32", ""), ("shape", 3, "int32", "repeated"),
                            ("dtype", 4, "string", ""), ("layer_id", 5, "int32", "")])
    _msg(fd, "ActivationRequest", [
        ("nonce", 1, "string", ""), ("activation", 2, ".dnetring.Activation", ""), ("timestamp", 3, "int64", ""),
        ("node_origin", 4, "string", ""), ("callback_url", 5, "string", ""), ("logprobs", 6, "bool", ""),
        ("top_logprobs", 7, "int32", ""), ("temperature", 8, "float", "optional"), ("top_p", 9, "float", "optional"),
        ("top_k", 10, "int32", "optional"), ("repetition_penalty", 11, "float", "optional"),
        ("min_p", 12, "float", "optional"), ("min_tokens_to_keep", 13, "int32", "optional")])
    _msg(fd, "ActivationResponse", [("success", 1, "bool", ""), ("message", 2, "string", ""), ("node_id", 3, "string", "")])
    _msg(fd, "ActivationFrame", [("request", 1, ".dnetring.ActivationRequest", ""), ("seq", 2, "uint64", ""),
                                 ("end_of_request", 3, "bool", "")])
    _msg(fd, "StreamAck", [("nonce", 1, "string", ""), ("seq", 2, "uint64", ""), ("accepted", 3, "bool", ""),
                           ("message", 4, "string", "")])
    _msg(fd, "HealthRequest", [("requester_id", 1, "string", "")])
    _msg(fd, "HealthResponse", [("healthy", 1, "bool", ""), ("node_id", 2, "string", ""),
                                ("assigned_layers", 3, "int32", "repeated"), ("queue_size", 4, "int32", ""),
                                ("active_requests", 5, "int32", "")])
    _msg(fd, "WeightRequest", [("weight_id", 1, "string", ""), ("layer_id", 2, "int32", ""), ("priority", 3, "int32", "")])
    _msg(fd, "ResetCacheRequest", [])
    _msg(fd, "ResetCacheResponse", [("success", 1, "bool", ""), ("message", 2, "string", "")])
    _msg(fd, "LatencyMeasureRequest", [("requester_id", 1, "string", ""), ("payload_size", 2, "int32", ""),
                                       ("dummy_data", 3, "bytes", ""), ("timestamp", 4, "int64", "")])
    _msg(fd, "LatencyMeasureResponse", [("success", 1, "bool", ""), ("message", 2, "string", ""),
                                        ("node_id", 3, "string", ""), ("timestamp", 4, "int64", "")])
    svc = fd.service.add()
    svc.name = "DnetRingService"
    for name, i, o, cs, ss in [("SendActivation", "ActivationRequest", "ActivationResponse", False, False),
                               ("HealthCheck", "HealthRequest", "HealthResponse", False, False),
                               ("ResetCache", "ResetCacheRequest", "ResetCacheResponse", False, False),
                               ("MeasureLatency", "LatencyMeasureRequest", "LatencyMeasureResponse", False, False),
                               ("StreamActivations", "ActivationFrame", "StreamAck", True, True)]:
        m = svc.method.add()
        m.name, m.input_type, m.output_type = name, ".dnetring." + i, ".dnetring." + o
        m.client_streaming, m.server_streaming = cs, ss
    return fd


def build_shard_api():
    fd = dp.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "shard_api_comm.proto", "shardapi", "proto3"
    _msg(fd, "FinalActivationRequest", [("nonce", 1, "string", ""), ("data", 2, "bytes", ""), ("batch_size", 3, "int32", ""),
                                        ("shape", 4, "int32", "repeated"), ("dtype", 5, "string", ""),
                                        ("layer_id", 6, "int32", ""), ("timestamp", 7, "int64", ""),
                                        ("node_origin", 8, "string", "")])
    _msg(fd, "FinalActivationResponse", [("success", 1, "bool", ""), ("message", 2, "string", ""), ("token_id", 3, "int32", "")])
    tr = _msg(fd, "TokenRequest", [("nonce", 1, "string", ""), ("token_id", 2, "int32", ""), ("timestamp", 3, "int64", ""),
                                   ("logprob", 4, "float", "")])
    # map<int32, float> top_logprobs = 5  == repeated nested TopLogprobsEntry{key=1,value=2} with map_entry
    ent = tr.nested_type.add()
    ent.name = "TopLogprobsEntry"
    ent.options.map_entry = True
    for fname, num, ft in (("key", 1, _T.TYPE_INT32), ("value", 2, _T.TYPE_FLOAT)):
        f = ent.field.add()
        f.name, f.number, f.label, f.type = fname, num, _T.LABEL_OPTIONAL, ft
    f = tr.field.add()
    f.name, f.number, f.label, f.type = "top_logprobs", 5, _T.LABEL_REPEATED, _T.TYPE_MESSAGE
    f.type_name = ".shardapi.TokenRequest.TopLogprobsEntry"
    _msg(fd, "TokenResponse", [("success", 1, "bool", ""), ("message", 2, "string", "")])
    _msg(fd, "RingError", [("nonce", 1, "string", ""), ("failed_node", 2, "string", ""), ("error_code", 3, "string", ""),
                           ("error", 4, "string", "")])
    svc = fd.service.add()
    svc.name = "ShardApiService"
    for name, i, o in [("SendFinalActivation", "FinalActivationRequest", "FinalActivationResponse"),
                       ("SendToken", "TokenRequest", "TokenResponse")]:
        m = svc.method.add()
        m.name, m.input_type, m.output_type = name, ".shardapi." + i, ".shardapi." + o
    return fd


_pool = descriptor_pool.DescriptorPool()
_files = {}


def messages(which: str) -> dict:
    if which not in _files:
        fd = build_ring() if which == "ring" else build_shard_api()
        _pool.Add(fd)
        fdesc = _pool.FindFileByName(fd.name)
        _files[which] = {name: message_factory.GetMessageClass(fdesc.message_types_by_name[name])
                         for name in fdesc.message_types_by_name}
    return _files[which]
