# Package src.retrieval -- only the file the B200 engine replaces lives here; in a deployment the
# reference's own classifier.py / orchestrator.py / responder.py stay beside it unchanged.
