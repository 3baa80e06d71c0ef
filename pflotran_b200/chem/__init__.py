"""Host-side chemistry setup: deck + thermodynamic database -> flat reaction tables."""
from .deck import read_deck, Deck, Chemistry, Constraint  # noqa: F401
from .basis import build_tables  # noqa: F401
from .tables import ReactionTables  # noqa: F401
