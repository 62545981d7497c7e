/*
 * JNI shim: com.justinblank.strings.gpu.NeedleNative -> include/needle_b200.h.
 *
 * Build where a JDK exists:  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../../../include \
 *     needle_jni.c -L../../../../needle_b200 -lneedle_b200 -o libneedle_jni.so
 * The build image of this repository has no JDK; tests/test_jni_compiles.py compiles this file against a compile-only
 * stub of jni.h (java/src/test/native/stub/jni.h, -Wall -Wextra -Werror), and every library call below is exercised through
 * the identical Python ctypes binding (needle_b200/_lib.py) in tests/.
 *
 * Rules kept here: no JNI critical region (GetStringCritical / GetPrimitiveArrayCritical) is held across a call into the
 * library - those calls block on the GPU and take a mutex, which JNI forbids inside a critical region and which can stall
 * the collector.  Strings and arrays are copied with Get*Region; batches travel in direct ByteBuffers, ideally the
 * page-locked ones of pinnedAlloc (ndl_host_alloc), which the library DMAs from directly (pageable buffers go through
 * its bounce ring).  Error codes map onto needle's exception classes exactly as include/needle_b200.h documents.
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "needle_b200.h"

static void throw_named(JNIEnv* env, const char* cls, const char* msg) {
  jclass c = (*env)->FindClass(env, cls);
  if (!c) {  /* (the class lookup itself left an exception pending - e.g. needle-types is not on the class path) */
    (*env)->ExceptionClear(env);
    c = (*env)->FindClass(env, "java/lang/RuntimeException");
  }
  if (c) (*env)->ThrowNew(env, c, msg);
}

static void throw_for(JNIEnv* env, int code) {
  const char* cls = "java/lang/RuntimeException"; /* NDL_ECUDA, NDL_ENCCL, NDL_EBLOB */
  if (code == NDL_ESYNTAX) cls = "com/justinblank/strings/PatternSyntaxException";                  /* RegexParser.java:91-97 */
  else if (code == NDL_ECOMPILE) cls = "com/justinblank/strings/PatternClassCompilationException";  /* DFACompiler.java:34-36 */
  else if (code == NDL_ETOOLARGE) cls = "java/lang/IllegalStateException";                           /* DFACompiler.java:76-83 */
  else if (code == NDL_EFLAGS || code == NDL_EINVAL) cls = "java/lang/IllegalArgumentException";     /* CompilerOptions.java:10-12 */
  else if (code == NDL_ENOMEM) cls = "java/lang/OutOfMemoryError";
  throw_named(env, cls, ndl_last_error());
}

static int null_arg(JNIEnv* env, const void* p, const char* what) {
  if (p) return 0;
  throw_named(env, "java/lang/NullPointerException", what);
  return 1;
}

JNIEXPORT jbyteArray JNICALL Java_com_justinblank_strings_gpu_NeedleNative_compile(JNIEnv* env, jclass k, jstring regex, jint flags) {
  (void)k;
  if (null_arg(env, regex, "regex string cannot be null")) return NULL;
  const jsize n = (*env)->GetStringLength(env, regex);
  jchar* chars = (jchar*)malloc(((size_t)n + 1) * sizeof(jchar)); /* UTF-16 code units, as ndl_compile wants */
  if (!chars) { throw_named(env, "java/lang/OutOfMemoryError", "regex copy"); return NULL; }
  (*env)->GetStringRegion(env, regex, 0, n, chars);
  uint8_t* blob = NULL;
  size_t len = 0;
  const int rc = ndl_compile((const uint16_t*)chars, (size_t)n, flags, &blob, &len);
  free(chars);
  if (rc != NDL_OK) { throw_for(env, rc); return NULL; }
  jbyteArray out = (*env)->NewByteArray(env, (jsize)len);
  if (out) (*env)->SetByteArrayRegion(env, out, 0, (jsize)len, (const jbyte*)blob);
  ndl_blob_free(blob);
  return out;
}

JNIEXPORT jlong JNICALL Java_com_justinblank_strings_gpu_NeedleNative_patternCreate(JNIEnv* env, jclass k, jbyteArray blob, jint device) {
  (void)k;
  if (null_arg(env, blob, "blob cannot be null")) return 0;
  const jsize len = (*env)->GetArrayLength(env, blob);
  jbyte* b = (jbyte*)malloc((size_t)len + 1);
  if (!b) { throw_named(env, "java/lang/OutOfMemoryError", "blob copy"); return 0; }
  (*env)->GetByteArrayRegion(env, blob, 0, len, b);
  ndl_pattern* p = NULL;
  const int rc = ndl_pattern_create((const uint8_t*)b, (size_t)len, device, &p); /* device -1: every visible GPU */
  free(b);
  if (rc != NDL_OK) { throw_for(env, rc); return 0; }
  return (jlong)(intptr_t)p;
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_patternDestroy(JNIEnv* env, jclass k, jlong h) {
  (void)env; (void)k;
  ndl_pattern_destroy((ndl_pattern*)(intptr_t)h);
}

/* Page-locked direct buffers for batches: ndl_host_alloc / ndl_host_free. */
JNIEXPORT jobject JNICALL Java_com_justinblank_strings_gpu_NeedleNative_pinnedAlloc(JNIEnv* env, jclass k, jlong bytes) {
  (void)k;
  if (bytes < 0) { throw_named(env, "java/lang/IllegalArgumentException", "negative size"); return NULL; }
  void* p = ndl_host_alloc((size_t)bytes);
  if (!p) { throw_named(env, "java/lang/OutOfMemoryError", ndl_last_error()); return NULL; }
  jobject buf = (*env)->NewDirectByteBuffer(env, p, bytes);
  if (!buf) ndl_host_free(p);
  return buf;
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_pinnedFree(JNIEnv* env, jclass k, jobject buffer) {
  (void)k;
  if (buffer) ndl_host_free((*env)->GetDirectBufferAddress(env, buffer));
}

/* Result arrays: the library writes into page-locked scratch, which is then copied into the Java arrays with Set*Region. */
typedef struct {
  uint8_t* matched;
  int32_t* start;
  int32_t* end;
} results_t;

static int results_alloc(JNIEnv* env, results_t* r, size_t n, int with_pos) {
  memset(r, 0, sizeof(*r));
  r->matched = (uint8_t*)ndl_host_alloc(n ? n : 1);
  if (with_pos) {
    r->start = (int32_t*)ndl_host_alloc((n ? n : 1) * sizeof(int32_t));
    r->end = (int32_t*)ndl_host_alloc((n ? n : 1) * sizeof(int32_t));
  }
  if (!r->matched || (with_pos && (!r->start || !r->end))) {
    ndl_host_free(r->matched); ndl_host_free(r->start); ndl_host_free(r->end);
    throw_named(env, "java/lang/OutOfMemoryError", "result staging");
    return -1;
  }
  return 0;
}

static void results_publish(JNIEnv* env, results_t* r, jsize n, jbyteArray matched, jintArray start, jintArray end, int ok) {
  if (ok) {
    (*env)->SetByteArrayRegion(env, matched, 0, n, (const jbyte*)r->matched);
    if (r->start && start) (*env)->SetIntArrayRegion(env, start, 0, n, (const jint*)r->start);
    if (r->end && end) (*env)->SetIntArrayRegion(env, end, 0, n, (const jint*)r->end);
  }
  ndl_host_free(r->matched); ndl_host_free(r->start); ndl_host_free(r->end);
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_matchBatch(JNIEnv* env, jclass k, jlong h, jint mode, jobject data,
                                                                                jobject offsets, jint n, jint charWidth,
                                                                                jbyteArray matched, jintArray start, jintArray end) {
  (void)k;
  if (null_arg(env, data, "data") || null_arg(env, offsets, "offsets") || null_arg(env, matched, "matched")) return;
  const void* d = (*env)->GetDirectBufferAddress(env, data);
  const uint64_t* o = (const uint64_t*)(*env)->GetDirectBufferAddress(env, offsets);
  if (!d || !o) { throw_named(env, "java/lang/IllegalArgumentException", "data and offsets must be direct buffers"); return; }
  results_t r;
  const int with_pos = mode == NDL_MODE_FIND;
  if (with_pos && (null_arg(env, start, "start") || null_arg(env, end, "end"))) return;
  if (results_alloc(env, &r, (size_t)n, with_pos)) return;
  const int rc = ndl_match_batch((ndl_pattern*)(intptr_t)h, mode, d, o, (uint64_t)n, charWidth, NULL, r.matched, r.start, r.end, NDL_MEM_HOST, NULL);
  results_publish(env, &r, n, matched, start, end, rc == NDL_OK);
  if (rc != NDL_OK) throw_for(env, rc);
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_matchLines(JNIEnv* env, jclass k, jlong h, jint mode, jobject data, jint n,
                                                                                jint lineChars, jint charWidth, jbyteArray matched,
                                                                                jintArray start, jintArray end) {
  (void)k;
  if (null_arg(env, data, "data") || null_arg(env, matched, "matched")) return;
  const void* d = (*env)->GetDirectBufferAddress(env, data);
  if (!d) { throw_named(env, "java/lang/IllegalArgumentException", "data must be a direct buffer"); return; }
  results_t r;
  const int with_pos = mode == NDL_MODE_FIND;
  if (with_pos && (null_arg(env, start, "start") || null_arg(env, end, "end"))) return;
  if (results_alloc(env, &r, (size_t)n, with_pos)) return;
  const int rc = ndl_match_lines((ndl_pattern*)(intptr_t)h, mode, d, (uint64_t)n, (uint64_t)lineChars, charWidth, r.matched, r.start, r.end,
                                 NDL_MEM_HOST, NULL);
  results_publish(env, &r, n, matched, start, end, rc == NDL_OK);
  if (rc != NDL_OK) throw_for(env, rc);
}

/* ndl_find_long on a direct buffer: ONE haystack (a memory-mapped file, a document), 64-bit offsets; on a pattern created for all
 * devices the library splits it across the GPUs.  Returns {matched, start, end}. */
JNIEXPORT jlongArray JNICALL Java_com_justinblank_strings_gpu_NeedleNative_findLong(JNIEnv* env, jclass k, jlong h, jobject data, jlong nChars,
                                                                                    jint charWidth, jlong from) {
  (void)k;
  if (null_arg(env, data, "data")) return NULL;
  const void* d = (*env)->GetDirectBufferAddress(env, data);
  const jlong cap = (*env)->GetDirectBufferCapacity(env, data);
  if (!d) { throw_named(env, "java/lang/IllegalArgumentException", "data must be a direct buffer"); return NULL; }
  if (nChars < 0 || (charWidth != 1 && charWidth != 2) || nChars * charWidth > cap) {
    throw_named(env, "java/lang/IllegalArgumentException", "nChars * charWidth exceeds the buffer");
    return NULL;
  }
  uint8_t matched = 0;
  int64_t start = -1, end = -1;
  const int rc = ndl_find_long((ndl_pattern*)(intptr_t)h, d, (uint64_t)nChars, charWidth, (int64_t)from, &matched, &start, &end, NDL_MEM_HOST, NULL);
  if (rc != NDL_OK) { throw_for(env, rc); return NULL; }
  jlongArray out = (*env)->NewLongArray(env, 3);
  if (!out) return NULL;
  const jlong v[3] = {matched, start, end};
  (*env)->SetLongArrayRegion(env, out, 0, 3, v);
  return out;
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_findAllBatch(JNIEnv* env, jclass k, jlong h, jobject data, jobject offsets,
                                                                                  jint n, jint charWidth, jintArray counts,
                                                                                  jobject matchOffsets, jintArray starts, jintArray ends) {
  (void)k;
  if (null_arg(env, data, "data") || null_arg(env, offsets, "offsets") || null_arg(env, counts, "counts")) return;
  const void* d = (*env)->GetDirectBufferAddress(env, data);
  const uint64_t* o = (const uint64_t*)(*env)->GetDirectBufferAddress(env, offsets);
  const uint64_t* mo = matchOffsets ? (const uint64_t*)(*env)->GetDirectBufferAddress(env, matchOffsets) : NULL;
  if (!d || !o || (matchOffsets && !mo)) { throw_named(env, "java/lang/IllegalArgumentException", "direct buffers required"); return; }
  if (mo && (null_arg(env, starts, "starts") || null_arg(env, ends, "ends"))) return;
  const size_t total = mo ? (size_t)(mo[n] - mo[0]) : 0;
  uint32_t* c = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
  int32_t* s = mo ? (int32_t*)malloc((total + 1) * sizeof(int32_t)) : NULL;
  int32_t* e = mo ? (int32_t*)malloc((total + 1) * sizeof(int32_t)) : NULL;
  if (!c || (mo && (!s || !e))) {
    free(c); free(s); free(e);
    throw_named(env, "java/lang/OutOfMemoryError", "result staging");
    return;
  }
  /* the C ABI indexes starts / ends by match_offsets[i] + k: hand it arrays whose element match_offsets[0] is ours */
  const int rc = ndl_find_all_batch((ndl_pattern*)(intptr_t)h, d, o, (uint64_t)n, charWidth, c, mo, mo ? s - mo[0] : NULL, mo ? e - mo[0] : NULL,
                                    NDL_MEM_HOST, NULL);
  if (rc == NDL_OK) {
    (*env)->SetIntArrayRegion(env, counts, 0, n, (const jint*)c);
    if (mo) {
      (*env)->SetIntArrayRegion(env, starts, (jsize)mo[0], (jsize)total, (const jint*)s);
      (*env)->SetIntArrayRegion(env, ends, (jsize)mo[0], (jsize)total, (const jint*)e);
    }
  }
  free(c); free(s); free(e);
  if (rc != NDL_OK) throw_for(env, rc);
}

JNIEXPORT jintArray JNICALL Java_com_justinblank_strings_gpu_NeedleNative_matchOne(JNIEnv* env, jclass k, jlong h, jint mode, jstring s, jint from) {
  (void)k;
  if (null_arg(env, s, "string cannot be null")) return NULL;
  const jsize n = (*env)->GetStringLength(env, s);
  jchar* chars = (jchar*)malloc(((size_t)n + 1) * sizeof(jchar));
  if (!chars) { throw_named(env, "java/lang/OutOfMemoryError", "string copy"); return NULL; }
  (*env)->GetStringRegion(env, s, 0, n, chars);
  uint64_t offsets[2] = {0, (uint64_t)n};
  int32_t f = from, st = -1, en = -1;
  uint8_t m = 0;
  const int rc = ndl_match_batch((ndl_pattern*)(intptr_t)h, mode, chars, offsets, 1, 2, &f, &m, &st, &en, NDL_MEM_HOST, NULL);
  free(chars);
  if (rc != NDL_OK) { throw_for(env, rc); return NULL; }
  const jint vals[3] = {m, st, en};
  jintArray out = (*env)->NewIntArray(env, 3);
  if (out) (*env)->SetIntArrayRegion(env, out, 0, 3, vals);
  return out;
}

/* findAllStrings: GetStringRegion straight into ONE page-locked UTF-16 staging buffer (ndl_host_alloc) + offsets, then one
 * ndl_match_batch - the strings are copied once, from the Java heap into memory the GPU's copy engine reads directly. */
JNIEXPORT jobject JNICALL Java_com_justinblank_strings_gpu_NeedleNative_findAllStrings(JNIEnv* env, jclass k, jlong h, jobjectArray hay) {
  (void)k;
  if (null_arg(env, hay, "haystacks cannot be null")) return NULL;
  const jsize n = (*env)->GetArrayLength(env, hay);
  uint64_t* offsets = (uint64_t*)ndl_host_alloc(((size_t)n + 1) * sizeof(uint64_t));
  if (!offsets) { throw_named(env, "java/lang/OutOfMemoryError", ndl_last_error()); return NULL; }
  uint64_t total = 0;
  offsets[0] = 0;
  for (jsize i = 0; i < n; i++) {
    jstring s = (jstring)(*env)->GetObjectArrayElement(env, hay, i);
    if (!s) {
      ndl_host_free(offsets);
      throw_named(env, "java/lang/NullPointerException", "haystacks must not contain null");
      return NULL;
    }
    total += (uint64_t)(*env)->GetStringLength(env, s);
    offsets[i + 1] = total;
    (*env)->DeleteLocalRef(env, s);
  }
  jchar* data = (jchar*)ndl_host_alloc((size_t)(total ? total : 1) * sizeof(jchar));
  jobject res = NULL;
  results_t r;
  memset(&r, 0, sizeof(r));
  if (!data) { throw_named(env, "java/lang/OutOfMemoryError", ndl_last_error()); goto done; }
  for (jsize i = 0; i < n; i++) {
    jstring s = (jstring)(*env)->GetObjectArrayElement(env, hay, i);
    (*env)->GetStringRegion(env, s, 0, (jsize)(offsets[i + 1] - offsets[i]), data + offsets[i]);
    (*env)->DeleteLocalRef(env, s);
  }
  {
    jclass rc_cls = (*env)->FindClass(env, "com/justinblank/strings/gpu/GpuPattern$BatchResult");
    if (!rc_cls) goto done; /* (exception pending) */
    jmethodID ctor = (*env)->GetMethodID(env, rc_cls, "<init>", "(I)V");
    jfieldID f_m = (*env)->GetFieldID(env, rc_cls, "matched", "[B"), f_s = (*env)->GetFieldID(env, rc_cls, "start", "[I"),
             f_e = (*env)->GetFieldID(env, rc_cls, "end", "[I");
    if (!ctor || !f_m || !f_s || !f_e) goto done;
    jobject obj = (*env)->NewObject(env, rc_cls, ctor, n);
    if (!obj) goto done;
    if (results_alloc(env, &r, (size_t)n, 1)) { memset(&r, 0, sizeof(r)); goto done; }
    const int rc = ndl_match_batch((ndl_pattern*)(intptr_t)h, NDL_MODE_FIND, data, offsets, (uint64_t)n, 2, NULL, r.matched, r.start, r.end,
                                   NDL_MEM_HOST, NULL);
    results_publish(env, &r, n, (jbyteArray)(*env)->GetObjectField(env, obj, f_m), (jintArray)(*env)->GetObjectField(env, obj, f_s),
                    (jintArray)(*env)->GetObjectField(env, obj, f_e), rc == NDL_OK);
    if (rc != NDL_OK) throw_for(env, rc);
    else res = obj;
  }
done:
  ndl_host_free(data);
  ndl_host_free(offsets);
  return res;
}
