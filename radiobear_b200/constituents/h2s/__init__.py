"""h2s absorption formalisms."""
