"""Plays the reference's known-answer scenarios (tests/golden/reference_kat.py) against a backend.

A backend exposes get_rate_limits(list_of_req_dicts) -> list_of_resp_dicts, now() and advance(ms) — the
frozen-clock shape of the reference's functional tests (clock.Freeze / clock.Advance).
"""
from golden import reference_kat as K


def play_scenario(backend, sc):
    for i, step in enumerate(sc["steps"]):
        req = dict(sc["req"])
        req.update(step["req"])
        now = backend.now()
        resp = backend.get_rate_limits([req])[0]
        ctx = f"{sc['name']} ({sc['cite']}) step {i}: req={req} resp={resp}"
        exp = step["expect"]
        assert resp["error"] == exp["error"], ctx
        assert resp["status"] == exp["status"], ctx
        if "remaining" in exp:
            assert resp["remaining"] == exp["remaining"], ctx
        want_limit = sc["limit"] if sc["limit"] is not None else req["limit"]
        assert resp["limit"] == want_limit, ctx
        if step["reset"]:
            assert K.RESET_PREDICATES[step["reset"]](resp, now), ctx
        backend.advance(step["sleep"])


def play_missing_fields(backend):
    for i, (req, err, status) in enumerate(K.MISSING_FIELDS):
        resp = backend.get_rate_limits([req])[0]
        assert resp["error"] == err, (i, resp)
        assert resp["status"] == status, (i, resp)
