"""co absorption formalisms."""
