"""Quirk Q1 fixture shared by the CPU and GPU tests: a bad-partial-key item whose signature VERIFIES, built from the reference's
own valid ceremony test_vectors/no_auth/finalization/report-1.json, so that prove_wrong_final_key_generation reaches
verify_expected_key / compute_pubkey_share (crates/dkg/src/verification.rs:399-420, 523-551) - a branch none of the reference's 17
wrong_final_key_generation vectors gets to.  The expected keys are SURVEY.md App. C3's golden values."""
import copy
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "tests", "golden", "reference_vectors", "no_auth", "finalization", "report-1.json")
BADREPORT = os.path.join(ROOT, "tests", "golden", "reference_vectors", "no_auth", "wrong_final_key_generation",
                         "badreport-1-gen-wrong-partial-pubkey.json")
# Horner over the final keys K_1..K_3 at id 1 / 2 / 3 (perpetrator index in base_hash-sorted order)
Q1_EXPECTED = ["959990bc5e49a5688f6f3b5188131bf0689b71953e55a990e1dffae8d083adcfeda42e967a93563a2ac499ddd3f2b06c",
               "8e2e678499cf9ee33c1b0f76d163a4434587377272c8babd5ecf6c76749ad220acae64d43b7020ad099520253db7be9e",
               "a4696b6a7c36ba4df67d2c0d7f508825dad9bd553750b7a079dd86df6d863d2f44657f452ae49670b833e0563db702d8"]
# the final keys themselves (App. C3): each equals the generation's partial_pubkey
FINAL_KEYS = ["a39ed53c850eecf70edaa9037057fe6e3a09909a16690e83c958bd2cc92b92a794c6bd0ac42b9a1c5e43169d71dd6e99",
              "8cc3a8f33d252ba45e2fb8e977259ce04645597ba50d37664bb8246374ca355c14553e52694a646502582eee1a45d539",
              "b9701cac5492592c2f4e1345ed34f26e6baaf7b2fb79f7203194bc493882b88e9120ba1fd2407f80f6fefab00b3886b4"]


def sorted_generations():
    rep = json.load(open(REPORT))["scenario"]
    return sorted(rep["generations"], key=lambda g: bytes.fromhex(g["base_hash"]))


def q1_item(perp):
    """no_auth BadPartialShareData accusing the generation at sorted index `perp` with its own, VALID, partial key + signature."""
    rep = json.load(open(REPORT))["scenario"]
    tmpl = json.load(open(BADREPORT))["scenario"]
    g = sorted_generations()[perp]
    item = {"settings": rep["settings"],
            "generations": [{"base_pubkeys": x["base_pubkeys"], "base_hash": x["base_hash"]} for x in rep["generations"]],
            "bad_partial": {"settings": rep["settings"], "data": copy.deepcopy(g), "commitment": tmpl["bad_partial"]["commitment"]}}
    return item
