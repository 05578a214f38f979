"""Two-tower retrieval task.  (The reference's retrieval/__init__.py:4-5 re-exports
FactorizedTopK and GCN, which import faiss at module load; both are outside the hot path --
SURVEY.md section 2 rows 2 -- so this package exports only the sbcnm task layers.)"""
from . import sbcnm  # noqa: F401
