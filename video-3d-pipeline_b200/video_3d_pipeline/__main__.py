"""`python -m video_3d_pipeline` runs the depth step's command line, like the reference's __main__.py:3-6."""
import sys

from . import depth

if __name__ == "__main__":
    sys.exit(depth.main())
