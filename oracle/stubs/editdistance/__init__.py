"""Stub for the un-vendored `editdistance` pip package (C++/Cython, unpinned).

Test infrastructure only. `editdistance.eval(a, b)` is plain unit-cost
Levenshtein distance on Python str, so any correct implementation is
result-identical. Used by /root/reference/barcode_graph.py:96,
243,315,379 when the unmodified reference is imported by oracle/ref_harness.py.
"""


def eval(a, b):  # noqa: A001 - name fixed by the package being stubbed
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


distance = eval
