/* TEST INFRASTRUCTURE ONLY (oracle/): the handful of OpenGL 3.3 entry points and enums the GLSL oracle needs,
 * resolved through the vendor's getProcAddress (no GL headers in this image; values from the Khronos registry). */
#ifndef GL_MIN_H
#define GL_MIN_H

#include <stddef.h>
#include <stdint.h>

typedef unsigned int GLuint, GLenum, GLbitfield;
typedef int GLint, GLsizei;
typedef unsigned char GLboolean, GLubyte;
typedef char GLchar;
typedef float GLfloat;
typedef ptrdiff_t GLsizeiptr, GLintptr;

#define GL_VENDOR 0x1F00
#define GL_RENDERER 0x1F01
#define GL_VERSION 0x1F02
#define GL_SHADING_LANGUAGE_VERSION 0x8B8C
#define GL_VERTEX_SHADER 0x8B31
#define GL_FRAGMENT_SHADER 0x8B30
#define GL_COMPILE_STATUS 0x8B81
#define GL_LINK_STATUS 0x8B82
#define GL_INTERLEAVED_ATTRIBS 0x8C8C
#define GL_ARRAY_BUFFER 0x8892
#define GL_TRANSFORM_FEEDBACK_BUFFER 0x8C8E
#define GL_STATIC_DRAW 0x88E4
#define GL_STATIC_READ 0x88E5
#define GL_FLOAT 0x1406
#define GL_POINTS 0x0000
#define GL_RASTERIZER_DISCARD 0x8C89
#define GL_TEXTURE_2D 0x0DE1
#define GL_TEXTURE0 0x84C0
#define GL_RGBA 0x1908
#define GL_RGBA8 0x8058
#define GL_RG 0x8227
#define GL_RG8 0x822B
#define GL_UNSIGNED_BYTE 0x1401
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_TEXTURE_WRAP_S 0x2802
#define GL_TEXTURE_WRAP_T 0x2803
#define GL_LINEAR 0x2601
#define GL_NEAREST 0x2600
#define GL_REPEAT 0x2901
#define GL_CLAMP_TO_EDGE 0x812F
#define GL_UNPACK_ALIGNMENT 0x0CF5
#define GL_TRANSFORM_FEEDBACK_PRIMITIVES_WRITTEN 0x8C88
#define GL_QUERY_RESULT 0x8866

typedef struct GlFns {
	const GLubyte *(*GetString)(GLenum);
	GLenum (*GetError)(void);
	GLuint (*CreateShader)(GLenum);
	void (*ShaderSource)(GLuint, GLsizei, const GLchar *const *, const GLint *);
	void (*CompileShader)(GLuint);
	void (*GetShaderiv)(GLuint, GLenum, GLint *);
	void (*GetShaderInfoLog)(GLuint, GLsizei, GLsizei *, GLchar *);
	GLuint (*CreateProgram)(void);
	void (*AttachShader)(GLuint, GLuint);
	void (*TransformFeedbackVaryings)(GLuint, GLsizei, const GLchar *const *, GLenum);
	void (*LinkProgram)(GLuint);
	void (*GetProgramiv)(GLuint, GLenum, GLint *);
	void (*GetProgramInfoLog)(GLuint, GLsizei, GLsizei *, GLchar *);
	void (*UseProgram)(GLuint);
	GLint (*GetUniformLocation)(GLuint, const GLchar *);
	void (*Uniform4fv)(GLint, GLsizei, const GLfloat *);
	void (*Uniform1i)(GLint, GLint);
	void (*GenVertexArrays)(GLsizei, GLuint *);
	void (*BindVertexArray)(GLuint);
	void (*GenBuffers)(GLsizei, GLuint *);
	void (*BindBuffer)(GLenum, GLuint);
	void (*BufferData)(GLenum, GLsizeiptr, const void *, GLenum);
	void (*GetBufferSubData)(GLenum, GLintptr, GLsizeiptr, void *);
	void (*BindBufferBase)(GLenum, GLuint, GLuint);
	void (*DeleteBuffers)(GLsizei, const GLuint *);
	void (*EnableVertexAttribArray)(GLuint);
	void (*VertexAttribPointer)(GLuint, GLint, GLenum, GLboolean, GLsizei, const void *);
	void (*Enable)(GLenum);
	void (*Disable)(GLenum);
	void (*BeginTransformFeedback)(GLenum);
	void (*EndTransformFeedback)(void);
	void (*DrawArrays)(GLenum, GLint, GLsizei);
	void (*Finish)(void);
	void (*GenTextures)(GLsizei, GLuint *);
	void (*BindTexture)(GLenum, GLuint);
	void (*ActiveTexture)(GLenum);
	void (*TexImage2D)(GLenum, GLint, GLint, GLsizei, GLsizei, GLint, GLenum, GLenum, const void *);
	void (*TexParameteri)(GLenum, GLenum, GLint);
	void (*PixelStorei)(GLenum, GLint);
	void (*GenQueries)(GLsizei, GLuint *);
	void (*BeginQuery)(GLenum, GLuint);
	void (*EndQuery)(GLenum);
	void (*GetQueryObjectuiv)(GLuint, GLenum, GLuint *);
} GlFns;

/* Resolves every entry point; returns NULL or the name of the first one the driver does not have. */
const char *gl_load(GlFns *gl, void *(*getProcAddress)(const char *));
/* Vertex shader only (rasterizer discarded), the listed varyings captured interleaved. 0 on failure (log filled). */
GLuint gl_build_tf_program(GlFns *gl, const char *vsSource, const char *const *varyings, int numVaryings, char *log, size_t logSize);
/* Draws `numVertices` points (gl_VertexID = 0..n-1) with attribute 0 (and 1 when attr1Offset >= 0) as float4 read from
 * `vertexData` at `stride` bytes per vertex, captures the varyings into `out`. 0 on success. */
int gl_run_tf(GlFns *gl, GLuint prog, const void *vertexData, size_t vertexBytes, int stride, const void *unused, int attr1Offset,
	int numVertices, void *out, size_t outBytes);

#endif
