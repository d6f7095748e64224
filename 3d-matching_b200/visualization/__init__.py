from .draw_registration_result import draw_registration_result

__all__ = ["draw_registration_result"]
