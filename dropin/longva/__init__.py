from streamchat_b200.longva import *  # noqa: F401,F403
