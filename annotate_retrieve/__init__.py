"""Import shims so that code written against the reference's module paths
(``annotate_retrieve.modeling_dense_passage_retrieval``, ``annotate_retrieve.modeling_iterative_rag``)
picks up the B200 implementation unchanged.  See INTEGRATION.md."""
