"""Drop-in counterparts of the reference's ``subgraph_isomorphism`` hot path (models + augmentations)."""
