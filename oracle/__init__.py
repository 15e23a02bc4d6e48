"""CPU oracle for the lossyless hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``lossyless_b200`` never does.

PARITY UNPINNED at the compressai / openai-CLIP boundary: neither package is vendored in
/root/reference nor installable offline, and the reference has no tests or golden
bitstreams.  The restatements below follow the published algorithms of
compressai==1.1.5 and clip==1.0 (pins: /root/reference/requirements/environment.yaml:98,105)
and are anchored on the reference's own call sites (cited per function) and on the few
reproducible known answers it ships (tests/test_oracle.py).
"""
