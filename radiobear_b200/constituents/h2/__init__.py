"""h2 absorption formalisms."""
