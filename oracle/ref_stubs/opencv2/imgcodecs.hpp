// Empty stand-in: /root/reference/src/registration.cpp:3 includes <opencv2/imgcodecs.hpp> but uses nothing from it.
// Only on the include path of `make -C oracle ref` when OpenCV's C++ headers are absent (OPENCV_STUB=1, the default).
#pragma once
