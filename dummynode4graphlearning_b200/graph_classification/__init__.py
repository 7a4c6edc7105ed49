"""Drop-in counterparts of the reference's ``graph_classification`` hot path (GIN / RGIN on PyG-style batches)."""
