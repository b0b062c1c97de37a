"""ph3 absorption formalisms."""
