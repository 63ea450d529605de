from streamchat_b200.longva.constants import *  # noqa: F401,F403
