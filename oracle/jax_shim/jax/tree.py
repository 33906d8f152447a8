from .tree_util import tree_map as map  # noqa: A001,F401
from .tree_util import tree_leaves as leaves  # noqa: F401
from .tree_util import tree_structure as structure  # noqa: F401
from .tree_util import tree_flatten as flatten  # noqa: F401
