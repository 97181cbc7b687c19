from .utils.extension import cpp  # same re-export as the reference's tetranerf/__init__.py:1
