"""`generate.v1` protobuf messages and the TextGenerationService gRPC surface, built at run time.

The reference generates `text_generation_server/pb/generate_pb2(_grpc).py` from /root/reference/proto/generate.proto
with grpc_tools at install time (server/Makefile:7-17); neither protoc nor grpc_tools exists in this image, so the same
descriptors are constructed here programmatically.  Field names, numbers, types, labels and proto3 `optional`
presence are identical to proto/generate.proto:5-224 (pinned by tests/test_pb.py against a fixture extracted from
that file), hence wire-compatible with the unchanged Rust router (router/client).

Use:  `from tgis_b200 import pb as generate_pb2` — `generate_pb2.Batch`, `.Request`, ... like the generated module.
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
PACKAGE = "generate.v1"
SERVICE = "TextGenerationService"

_T = {"float": _F.TYPE_FLOAT, "uint32": _F.TYPE_UINT32, "uint64": _F.TYPE_UINT64, "bool": _F.TYPE_BOOL,
      "string": _F.TYPE_STRING}

# message -> [(name, number, type, label)]; label: "" | "repeated" | "optional"; type: scalar | message | enum name
SPEC = {
    "HealthRequest": [], "HealthResponse": [], "ServiceDiscoveryRequest": [],
    "ServiceDiscoveryResponse": [("urls", 1, "string", "repeated")],
    "ClearCacheRequest": [], "ClearCacheResponse": [], "ModelInfoRequest": [],
    "MemoryScalingModel": [("prefill_linear_coef0", 1, "float", ""), ("prefill_quadratic_coef0", 2, "float", ""),
                           ("prefill_quadratic_coef1", 3, "float", ""), ("nexttoken_linear_coef0", 4, "float", ""),
                           ("nexttoken_linear_coef1", 5, "float", ""), ("weight_limit", 6, "uint64", "")],
    "ModelInfoResponse": [("model_type", 1, "ModelInfoResponse.ModelType", ""), ("eos_token", 2, "uint32", ""),
                          ("batch_padding", 3, "bool", ""), ("memory_scaling_model", 4, "MemoryScalingModel", "")],
    "NextTokenChooserParameters": [("temperature", 1, "float", ""), ("top_k", 2, "uint32", ""), ("top_p", 3, "float", ""),
                                   ("typical_p", 4, "float", ""), ("min_new_tokens", 100, "uint32", ""),
                                   ("seed", 101, "uint64", "optional"), ("repetition_penalty", 102, "float", "optional"),
                                   ("length_penalty", 103, "NextTokenChooserParameters.LengthPenalty", "optional")],
    "RequestedDetails": [("input_toks", 1, "bool", ""), ("logprobs", 2, "bool", ""), ("ranks", 3, "bool", ""),
                         ("top_n_toks", 4, "uint32", "")],
    "Request": [("id", 1, "uint64", ""), ("prefix_id", 2, "string", ""), ("inputs", 3, "string", ""),
                ("input_length", 4, "uint32", ""), ("truncate", 5, "bool", ""), ("max_output_length", 6, "uint32", ""),
                ("parameters", 7, "NextTokenChooserParameters", ""), ("stream_response", 100, "bool", ""),
                ("details", 101, "RequestedDetails", "")],
    "StopSequence": [("tokens", 1, "uint32", "repeated")],
    "Batch": [("id", 1, "uint64", ""), ("requests", 2, "Request", "repeated"), ("total_tokens", 3, "uint32", "")],
    "TopToken": [("token_id", 1, "uint32", ""), ("logprob", 2, "float", "")],
    "Token": [("request_id", 1, "uint64", ""), ("token_id", 2, "uint32", ""), ("logprob", 3, "float", ""),
              ("rank", 4, "uint32", ""), ("top_tokens", 5, "TopToken", "repeated")],
    "GenerateError": [("request_id", 1, "uint64", ""), ("message", 2, "string", "")],
    "InputTokens": [("request_id", 1, "uint64", ""), ("tokens", 2, "Token", "repeated")],
    "PrefillRequest": [("batch", 1, "Batch", ""), ("to_prune", 2, "CachedBatch", "repeated")],
    "GenerateResult": [("output_tokens", 1, "Token", "repeated"), ("errors", 2, "GenerateError", "repeated"),
                       ("batch_id", 3, "uint64", ""), ("forward_time_ns", 4, "uint64", "")],
    "PrefillResponse": [("result", 1, "GenerateResult", ""), ("input_tokens", 2, "InputTokens", "repeated")],
    "RequestsStatus": [("completed_ids", 3, "uint64", "repeated")],
    "CachedBatch": [("batch_id", 1, "uint64", ""), ("status", 2, "RequestsStatus", "optional")],
    "NextTokenRequest": [("batches", 1, "CachedBatch", "repeated")],
    "NextTokenResponse": [("result", 1, "GenerateResult", "optional")],
    "PruneBatchRequest": [("batch", 1, "CachedBatch", "")],
    "PruneBatchResponse": [("batch_id", 1, "uint64", "optional")],
    "PrefixLookupRequest": [("prefix_id", 1, "string", "")],
    "PrefixLookupResponse": [("prefix_length", 1, "uint32", "")],
}
NESTED = {
    "NextTokenChooserParameters": {"LengthPenalty": [("start_index", 1, "uint32", ""), ("decay_factor", 2, "float", "")]},
}
ENUMS = {"ModelInfoResponse": {"ModelType": [("CAUSAL_LM", 0), ("SEQ2SEQ_LM", 1)]}}
METHODS = [("ServiceDiscovery", "ServiceDiscoveryRequest", "ServiceDiscoveryResponse"),
           ("ClearCache", "ClearCacheRequest", "ClearCacheResponse"), ("ModelInfo", "ModelInfoRequest", "ModelInfoResponse"),
           ("Prefill", "PrefillRequest", "PrefillResponse"), ("NextToken", "NextTokenRequest", "NextTokenResponse"),
           ("PruneBatch", "PruneBatchRequest", "PruneBatchResponse"),
           ("PrefixLookup", "PrefixLookupRequest", "PrefixLookupResponse"), ("Health", "HealthRequest", "HealthResponse")]


def _add_fields(msg: descriptor_pb2.DescriptorProto, fields, enum_names):
    n_oneof = 0
    for name, number, typ, label in fields:
        f = msg.field.add()
        f.name, f.number = name, number
        f.json_name = name
        f.label = _F.LABEL_REPEATED if label == "repeated" else _F.LABEL_OPTIONAL
        if typ in _T:
            f.type = _T[typ]
        elif typ in enum_names:
            f.type = _F.TYPE_ENUM
            f.type_name = f".{PACKAGE}.{typ}"
        else:
            f.type = _F.TYPE_MESSAGE
            f.type_name = f".{PACKAGE}.{typ}"
        if label == "optional":  # proto3 explicit presence = synthetic oneof
            msg.oneof_decl.add().name = f"_{name}"
            f.oneof_index = n_oneof
            f.proto3_optional = True
            n_oneof += 1


def _build_file() -> descriptor_pb2.FileDescriptorProto:
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "generate.proto"
    fd.package = PACKAGE
    fd.syntax = "proto3"
    enum_names = {f"{m}.{e}" for m, es in ENUMS.items() for e in es}
    for mname, fields in SPEC.items():
        msg = fd.message_type.add()
        msg.name = mname
        for ename, values in ENUMS.get(mname, {}).items():
            e = msg.enum_type.add()
            e.name = ename
            for vname, vnum in values:
                v = e.value.add()
                v.name, v.number = vname, vnum
        for nname, nfields in NESTED.get(mname, {}).items():
            nm = msg.nested_type.add()
            nm.name = nname
            _add_fields(nm, nfields, enum_names)
        _add_fields(msg, fields, enum_names)
    svc = fd.service.add()
    svc.name = SERVICE
    for mname, req, res in METHODS:
        m = svc.method.add()
        m.name, m.input_type, m.output_type = mname, f".{PACKAGE}.{req}", f".{PACKAGE}.{res}"
    return fd


FILE_DESCRIPTOR_PROTO = _build_file()
_pool = descriptor_pool.DescriptorPool()
_pool.AddSerializedFile(FILE_DESCRIPTOR_PROTO.SerializeToString())
DESCRIPTOR = _pool.FindFileByName("generate.proto")

for _name in SPEC:
    globals()[_name] = message_factory.GetMessageClass(DESCRIPTOR.message_types_by_name[_name])

SERVICE_FULL_NAME = f"{PACKAGE}.{SERVICE}"


# ---------------------------------------------------------------------------------------------- gRPC glue
def add_TextGenerationServiceServicer_to_server(servicer, server):
    """Same role as generate_pb2_grpc.add_TextGenerationServiceServicer_to_server (used at server.py:430-435)."""
    import grpc

    handlers = {}
    for mname, req, res in METHODS:
        handlers[mname] = grpc.unary_unary_rpc_method_handler(
            getattr(servicer, mname),
            request_deserializer=globals()[req].FromString,
            response_serializer=globals()[res].SerializeToString,
        )
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE_FULL_NAME, handlers),))


class TextGenerationServiceStub:
    """Client stub (what the Rust `ShardedClient` is to the router); used by tests and bench to drive a shard."""

    def __init__(self, channel):
        for mname, req, res in METHODS:
            setattr(self, mname, channel.unary_unary(f"/{SERVICE_FULL_NAME}/{mname}",
                                                     request_serializer=globals()[req].SerializeToString,
                                                     response_deserializer=globals()[res].FromString))
