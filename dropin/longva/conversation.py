from streamchat_b200.longva.conversation import Conversation, SeparatorStyle, conv_templates, default_conversation  # noqa: F401
