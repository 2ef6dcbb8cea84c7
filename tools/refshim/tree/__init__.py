from jax.tree_util import tree_map as map_structure  # noqa: F401
