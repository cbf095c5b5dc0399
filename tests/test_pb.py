"""CPU tests: the run-time `generate.v1` descriptors are field-for-field the reference's proto/generate.proto
(fixture tests/golden/generate_proto_fields.json extracted from that file by tests/golden/make_golden.py)."""
import json
import os

import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pb():
    import tgis_b200  # noqa: F401
    from tgis_b200 import pb as _pb
    return _pb


def test_messages_fields_numbers_types_labels_match_the_proto(pb):
    ref = json.load(open(os.path.join(G, "generate_proto_fields.json")))
    from google.protobuf import descriptor as d
    tname = {d.FieldDescriptor.TYPE_FLOAT: "float", d.FieldDescriptor.TYPE_UINT32: "uint32", d.FieldDescriptor.TYPE_UINT64: "uint64",
             d.FieldDescriptor.TYPE_BOOL: "bool", d.FieldDescriptor.TYPE_STRING: "string"}
    msgs = {k: v for k, v in ref.items() if not k.startswith("__")}
    assert msgs, "fixture empty"

    def find(name):
        parts = name.split(".")
        m = pb.DESCRIPTOR.message_types_by_name[parts[0]]
        for p in parts[1:]:
            m = m.nested_types_by_name[p]
        return m

    seen = 0
    for mname, fields in msgs.items():
        m = find(mname)
        assert len(m.fields) == len(fields), mname
        for fname, number, typ, label in fields:
            f = m.fields_by_name[fname]
            assert f.number == number, (mname, fname)
            if f.type in tname:
                assert tname[f.type] == typ, (mname, fname)
            elif f.type == d.FieldDescriptor.TYPE_MESSAGE:
                assert f.message_type.name == typ.split(".")[-1], (mname, fname)
            else:
                assert f.enum_type.name == typ.split(".")[-1], (mname, fname)
            assert (f.label == d.FieldDescriptor.LABEL_REPEATED) == (label == "repeated"), (mname, fname)
            assert bool(f.has_presence and f.type != d.FieldDescriptor.TYPE_MESSAGE) == (label == "optional") or \
                f.type == d.FieldDescriptor.TYPE_MESSAGE, (mname, fname)
            seen += 1
    assert seen >= 60
    # every top-level message of the proto exists here and vice versa
    top = {k for k in msgs if "." not in k}
    assert top == set(pb.DESCRIPTOR.message_types_by_name), top ^ set(pb.DESCRIPTOR.message_types_by_name)
    rpcs = {tuple(r) for r in ref["__rpc__"]}
    assert rpcs == set(pb.METHODS)
    assert ref["__enum__.ModelInfoResponse.ModelType"] == [["CAUSAL_LM", 0], ["SEQ2SEQ_LM", 1]]


def test_wire_round_trip_and_optional_presence(pb):
    r = pb.Request(id=7, inputs="test test", input_length=2, truncate=True, max_output_length=5,
                   parameters=pb.NextTokenChooserParameters(temperature=0.0, top_p=1.0, min_new_tokens=5, seed=9),
                   details=pb.RequestedDetails(logprobs=True, top_n_toks=3))
    b = pb.Batch(id=3, requests=[r], total_tokens=7)
    b2 = pb.Batch.FromString(b.SerializeToString())
    assert b2 == b and b2.requests[0].parameters.HasField("seed") and not b2.requests[0].parameters.HasField("repetition_penalty")
    # RequestsStatus.completed_ids is field number 3 (proto/generate.proto:184-187)
    st = pb.RequestsStatus(completed_ids=[5])
    assert st.SerializeToString()[0] >> 3 == 3
    cb = pb.CachedBatch(batch_id=1)
    assert not cb.HasField("status")  # "status absent" = whole batch finished (server.py:191-199)
    assert not pb.NextTokenResponse().HasField("result")
    assert pb.SERVICE_FULL_NAME == "generate.v1.TextGenerationService"
