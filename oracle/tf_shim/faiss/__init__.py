"""Empty stand-in so `import faiss` at the top of the reference's factorized_top_k.py succeeds
(the Faiss index class is outside the hot path and never instantiated by the golden script)."""
