"""fruit_nerf/components/field_heads.py:29-40 -- SemanticFieldHead (a Linear in -> num_classes).
The parameter holder lives next to the field; re-exported here under the reference's module path."""
from ..fruit_field import SemanticFieldHead  # noqa: F401
