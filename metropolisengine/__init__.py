"""Drop-in import name of the reference package (``import metropolisengine as me``, reference README.md:15,33;
package surface metropolisengine/__init__.py:1 = the one class).  Everything lives in ``metropolisengine_b200``."""
from metropolisengine_b200 import *  # noqa: F401,F403
from metropolisengine_b200 import MetropolisEngine, __all__  # noqa: F401
