/* Build shim (oracle/, test infrastructure): the reference's Visualizer.h includes GL/glew.h for a handful of type
 * names.  No OpenGL exists in this environment and none is called: the viewer is replaced by gdel_visualizer_stub.cu. */
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef float GLfloat;
