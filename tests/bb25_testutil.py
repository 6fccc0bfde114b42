"""Shared helpers for the bb25 test-suite."""


def case_scores(arrays, meta):
    """CSC dict of a golden scorer case."""
    p = meta["prefix"]
    return {
        "data": arrays[p + "data"], "indices": arrays[p + "indices"],
        "indptr": arrays[p + "indptr"], "doc_len": arrays[p + "doc_len"],
        "num_docs": meta["num_docs"], "avgdl": meta["avgdl"],
    }
