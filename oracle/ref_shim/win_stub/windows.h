// Stand-in for <windows.h>: /root/reference/src/octree/Rle4.cpp:2,14 needs it only for MessageBox.
#pragma once
#include <stdio.h>
#define MessageBox(a, b, c, d) fprintf(stderr, "%s: %s\n", (c), (b))
