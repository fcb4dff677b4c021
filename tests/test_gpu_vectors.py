"""-m gpu: drive the reference's own test_vectors/ and examples/ (data fixtures under
tests/golden/reference_vectors) through the GPU-backed `execute` front end and compare the process
exit code with the vector's expected_exit_code (the reference's pass criterion, script/run.sh:78-82)
and the reached status with the table reproduced by the independent restatement
(tests/golden/vector_outcomes.json)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = os.path.join(ROOT, "tests", "golden", "reference_vectors")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "vector_outcomes.json")))
TYPE = {"share": "bad-share", "finalization": "finalization", "wrong_final_key_generation": "bad-partial-key",
        "bad_encrypted_share": "bad-encrypted-share"}


@pytest.mark.parametrize("entry", GOLD["vectors"], ids=lambda e: e["file"])
def test_reference_vector(verifier, entry):
    mode, kind, _ = entry["file"].split("/")
    j = json.load(open(os.path.join(VEC, entry["file"])))
    code, status, msg = verifier.execute(TYPE[kind], json.dumps(j["scenario"]), auth=(mode == "auth"))
    assert code == j["params"]["expected_exit_code"], (status, msg)
    assert status == entry["status"], (status, entry["status_name"], msg)


@pytest.mark.parametrize("entry", GOLD["examples"], ids=lambda e: f'{e["file"]}-auth{e["auth"]}')
def test_reference_example(verifier, entry):
    kind = {"dvt_bad_share.json": "bad-share", "finalization_test.json": "finalization", "bad_partial_key.json": "bad-partial-key"}[entry["file"]]
    text = open(os.path.join(VEC, "examples", entry["file"])).read()
    code, status, msg = verifier.execute(kind, text, auth=entry["auth"], bls_identity=True)
    assert (code, status) == (entry["exit_code"], entry["status"]), msg


def test_malformed_inputs(verifier):
    assert verifier.execute("bad-share", "{")[0:2] == (1, 255)
    assert verifier.execute("finalization", '{"settings": {"n": 300, "k": 2, "gen_id": "00"}}')[0:2] == (1, 255)
    j = json.load(open(os.path.join(VEC, "no_auth", "share", "seeds-commitment-from-2-to-1.json")))["scenario"]
    j["initial_commitment"]["base_pubkeys"][0] = j["initial_commitment"]["base_pubkeys"][0][:-2]  # wrong hex length
    assert verifier.execute("bad-share", json.dumps(j))[0:2] == (1, 255)
