"""aicam-b200: B200-native YOLOv8 + DeepSORT hot path behind the AI-Camera class interfaces."""
__version__ = "0.1.0"
