/* TEST INFRASTRUCTURE ONLY (oracle/): stage-1 check that a GL 3.3 core context comes up on the GPU box through
 * egl_direct and that a vertex shader runs with transform feedback. Prints what it finds; exit 0 = all good. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "egl_direct.h"
#include "gl_min.h"

int main(int argc, char **argv)
{
	EglDirect e;
	if (egl_direct_open(&e, argc > 1 ? argv[1] : NULL, 1) != 0) { fprintf(stderr, "egl_direct_open: %s\n", e.error); return 2; }
	GlFns gl;
	const char *missing = gl_load(&gl, e.getProcAddress);
	if (missing) { fprintf(stderr, "GL entry point missing: %s\n", missing); return 3; }
	printf("GL_VENDOR   %s\nGL_RENDERER %s\nGL_VERSION  %s\nGLSL        %s\n", (const char*)gl.GetString(GL_VENDOR), (const char*)gl.GetString(GL_RENDERER),
		(const char*)gl.GetString(GL_VERSION), (const char*)gl.GetString(GL_SHADING_LANGUAGE_VERSION));
	const char *src = "#version 330\nlayout(location=0) in vec4 a; out vec4 o; void main() { o = a * 2.0 + vec4(float(gl_VertexID)); gl_Position = vec4(0.0); }\n";
	const char *varyings[1] = { "o" };
	char log[4096];
	GLuint prog = gl_build_tf_program(&gl, src, varyings, 1, log, sizeof(log));
	if (!prog) { fprintf(stderr, "shader: %s\n", log); return 4; }
	float in[16], out[16];
	for (int i = 0; i < 16; i++) in[i] = (float)i * 0.5f;
	if (gl_run_tf(&gl, prog, in, sizeof(in), 16, NULL, 0, 4, out, sizeof(out)) != 0) { fprintf(stderr, "transform feedback failed, glGetError 0x%x\n", gl.GetError()); return 5; }
	int ok = 1;
	for (int i = 0; i < 16; i++) { float want = in[i] * 2.0f + (float)(i / 4); if (out[i] != want) ok = 0; }
	printf("transform feedback: %s (out[5] = %g)\n", ok ? "ok" : "MISMATCH", out[5]);
	egl_direct_close(&e);
	return ok ? 0 : 6;
}
