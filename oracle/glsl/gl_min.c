/* TEST INFRASTRUCTURE ONLY (oracle/): see gl_min.h. */
#include "gl_min.h"

#include <stdio.h>
#include <string.h>

const char *gl_load(GlFns *gl, void *(*gpa)(const char *))
{
#define L(name) do { *(void**)&gl->name = gpa("gl" #name); if (!gl->name) return "gl" #name; } while (0)
	L(GetString); L(GetError); L(CreateShader); L(ShaderSource); L(CompileShader); L(GetShaderiv); L(GetShaderInfoLog);
	L(CreateProgram); L(AttachShader); L(TransformFeedbackVaryings); L(LinkProgram); L(GetProgramiv); L(GetProgramInfoLog);
	L(UseProgram); L(GetUniformLocation); L(Uniform4fv); L(Uniform1i); L(GenVertexArrays); L(BindVertexArray); L(GenBuffers);
	L(BindBuffer); L(BufferData); L(GetBufferSubData); L(BindBufferBase); L(DeleteBuffers); L(EnableVertexAttribArray);
	L(VertexAttribPointer); L(Enable); L(Disable); L(BeginTransformFeedback); L(EndTransformFeedback); L(DrawArrays); L(Finish);
	L(GenTextures); L(BindTexture); L(ActiveTexture); L(TexImage2D); L(TexParameteri); L(PixelStorei);
	L(GenQueries); L(BeginQuery); L(EndQuery); L(GetQueryObjectuiv);
#undef L
	return NULL;
}

GLuint gl_build_tf_program(GlFns *gl, const char *vsSource, const char *const *varyings, int numVaryings, char *log, size_t logSize)
{
	GLint ok = 0;
	log[0] = 0;
	GLuint vs = gl->CreateShader(GL_VERTEX_SHADER);
	gl->ShaderSource(vs, 1, &vsSource, NULL);
	gl->CompileShader(vs);
	gl->GetShaderiv(vs, GL_COMPILE_STATUS, &ok);
	if (!ok) { gl->GetShaderInfoLog(vs, (GLsizei)logSize, NULL, log); return 0; }
	/* a core-profile program needs no fragment stage while rasterization is discarded */
	GLuint prog = gl->CreateProgram();
	gl->AttachShader(prog, vs);
	gl->TransformFeedbackVaryings(prog, numVaryings, varyings, GL_INTERLEAVED_ATTRIBS);
	gl->LinkProgram(prog);
	gl->GetProgramiv(prog, GL_LINK_STATUS, &ok);
	if (!ok) { gl->GetProgramInfoLog(prog, (GLsizei)logSize, NULL, log); return 0; }
	return prog;
}

int gl_run_tf(GlFns *gl, GLuint prog, const void *vertexData, size_t vertexBytes, int stride, const void *unused, int attr1Offset,
	int numVertices, void *out, size_t outBytes)
{
	(void)unused;
	GLuint vao = 0, vbo = 0, tfb = 0, query = 0, written = 0;
	gl->GenVertexArrays(1, &vao);
	gl->BindVertexArray(vao);
	gl->GenBuffers(1, &vbo);
	gl->BindBuffer(GL_ARRAY_BUFFER, vbo);
	gl->BufferData(GL_ARRAY_BUFFER, (GLsizeiptr)vertexBytes, vertexData, GL_STATIC_DRAW);
	gl->EnableVertexAttribArray(0);
	gl->VertexAttribPointer(0, 4, GL_FLOAT, 0, stride, (const void*)0);
	if (attr1Offset > 0) {
		gl->EnableVertexAttribArray(1);
		gl->VertexAttribPointer(1, 4, GL_FLOAT, 0, stride, (const void*)(intptr_t)attr1Offset);
	}
	gl->GenBuffers(1, &tfb);
	gl->BindBuffer(GL_TRANSFORM_FEEDBACK_BUFFER, tfb);
	gl->BufferData(GL_TRANSFORM_FEEDBACK_BUFFER, (GLsizeiptr)outBytes, NULL, GL_STATIC_READ);
	gl->BindBufferBase(GL_TRANSFORM_FEEDBACK_BUFFER, 0, tfb);
	gl->UseProgram(prog);
	gl->Enable(GL_RASTERIZER_DISCARD);
	gl->GenQueries(1, &query);
	gl->BeginQuery(GL_TRANSFORM_FEEDBACK_PRIMITIVES_WRITTEN, query);
	gl->BeginTransformFeedback(GL_POINTS);
	gl->DrawArrays(GL_POINTS, 0, numVertices);
	gl->EndTransformFeedback();
	gl->EndQuery(GL_TRANSFORM_FEEDBACK_PRIMITIVES_WRITTEN);
	gl->Finish();
	gl->GetQueryObjectuiv(query, GL_QUERY_RESULT, &written);
	gl->GetBufferSubData(GL_TRANSFORM_FEEDBACK_BUFFER, 0, (GLsizeiptr)outBytes, out);
	gl->Disable(GL_RASTERIZER_DISCARD);
	gl->DeleteBuffers(1, &vbo);
	gl->DeleteBuffers(1, &tfb);
	if (gl->GetError() != 0) return -1;
	return written == (GLuint)numVertices ? 0 : -2;
}
