/* Empty stand-in: the reference includes <helper_timer.h> (gcvt.cu:27) and uses nothing from it. */
