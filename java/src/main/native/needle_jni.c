/*
 * JNI shim: com.justinblank.strings.gpu.NeedleNative -> include/needle_b200.h.
 *
 * Built only where a JDK exists:  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux \
 *     -I../../../../include needle_jni.c -L../../../../needle_b200 -lneedle_b200 -o libneedle_jni.so
 * The build image of this repository has no jni.h, so this file is not part of the automated build; it is a
 * mechanical wrapper and every call below is exercised through the identical Python ctypes binding
 * (needle_b200/_lib.py) in tests/.
 */
#include <jni.h>
#include <stdlib.h>
#include <string.h>

#include "needle_b200.h"

static void throw_for(JNIEnv* env, int code) {
  const char* cls = "java/lang/RuntimeException";
  if (code == NDL_ESYNTAX) cls = "com/justinblank/strings/PatternSyntaxException";
  else if (code == NDL_ETOOLARGE) cls = "java/lang/IllegalStateException";
  else if (code == NDL_EFLAGS || code == NDL_EINVAL) cls = "java/lang/IllegalArgumentException";
  else if (code == NDL_ENOMEM) cls = "java/lang/OutOfMemoryError";
  jclass c = (*env)->FindClass(env, cls);
  if (c) (*env)->ThrowNew(env, c, ndl_last_error());
}

JNIEXPORT jbyteArray JNICALL Java_com_justinblank_strings_gpu_NeedleNative_compile(JNIEnv* env, jclass k, jstring regex, jint flags) {
  (void)k;
  jsize n = (*env)->GetStringLength(env, regex);
  const jchar* chars = (*env)->GetStringCritical(env, regex, NULL);   /* UTF-16 code units, as ndl_compile wants */
  uint8_t* blob = NULL;
  size_t len = 0;
  int rc = ndl_compile((const uint16_t*)chars, (size_t)n, flags, &blob, &len);
  (*env)->ReleaseStringCritical(env, regex, chars);
  if (rc != NDL_OK) { throw_for(env, rc); return NULL; }
  jbyteArray out = (*env)->NewByteArray(env, (jsize)len);
  if (out) (*env)->SetByteArrayRegion(env, out, 0, (jsize)len, (const jbyte*)blob);
  ndl_blob_free(blob);
  return out;
}

JNIEXPORT jlong JNICALL Java_com_justinblank_strings_gpu_NeedleNative_patternCreate(JNIEnv* env, jclass k, jbyteArray blob, jint device) {
  (void)k;
  jsize len = (*env)->GetArrayLength(env, blob);
  jbyte* b = (*env)->GetByteArrayElements(env, blob, NULL);
  ndl_pattern* p = NULL;
  int rc = ndl_pattern_create((const uint8_t*)b, (size_t)len, device, &p);
  (*env)->ReleaseByteArrayElements(env, blob, b, JNI_ABORT);
  if (rc != NDL_OK) { throw_for(env, rc); return 0; }
  return (jlong)(intptr_t)p;
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_patternDestroy(JNIEnv* env, jclass k, jlong h) {
  (void)env; (void)k;
  ndl_pattern_destroy((ndl_pattern*)(intptr_t)h);
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_matchBatch(JNIEnv* env, jclass k, jlong h, jint mode, jobject data,
                                                                                jobject offsets, jint n, jint charWidth,
                                                                                jbyteArray matched, jintArray start, jintArray end) {
  (void)k;
  const void* d = (*env)->GetDirectBufferAddress(env, data);
  const uint64_t* o = (const uint64_t*)(*env)->GetDirectBufferAddress(env, offsets);
  jbyte* m = (*env)->GetPrimitiveArrayCritical(env, matched, NULL);
  jint* s = (*env)->GetPrimitiveArrayCritical(env, start, NULL);
  jint* e = (*env)->GetPrimitiveArrayCritical(env, end, NULL);
  int rc = ndl_match_batch((ndl_pattern*)(intptr_t)h, mode, d, o, (uint64_t)n, charWidth, NULL, (uint8_t*)m, (int32_t*)s, (int32_t*)e,
                           NDL_MEM_HOST, NULL);
  (*env)->ReleasePrimitiveArrayCritical(env, end, e, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, start, s, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, matched, m, 0);
  if (rc != NDL_OK) throw_for(env, rc);
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_matchLines(JNIEnv* env, jclass k, jlong h, jint mode, jobject data, jint n,
                                                                                jint lineChars, jint charWidth, jbyteArray matched,
                                                                                jintArray start, jintArray end) {
  (void)k;
  const void* d = (*env)->GetDirectBufferAddress(env, data);
  jbyte* m = (*env)->GetPrimitiveArrayCritical(env, matched, NULL);
  jint* s = (*env)->GetPrimitiveArrayCritical(env, start, NULL);
  jint* e = (*env)->GetPrimitiveArrayCritical(env, end, NULL);
  int rc = ndl_match_lines((ndl_pattern*)(intptr_t)h, mode, d, (uint64_t)n, (uint64_t)lineChars, charWidth, (uint8_t*)m, (int32_t*)s,
                           (int32_t*)e, NDL_MEM_HOST, NULL);
  (*env)->ReleasePrimitiveArrayCritical(env, end, e, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, start, s, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, matched, m, 0);
  if (rc != NDL_OK) throw_for(env, rc);
}

JNIEXPORT void JNICALL Java_com_justinblank_strings_gpu_NeedleNative_findAllBatch(JNIEnv* env, jclass k, jlong h, jobject data, jobject offsets,
                                                                                  jint n, jint charWidth, jintArray counts,
                                                                                  jobject matchOffsets, jintArray starts, jintArray ends) {
  (void)k;
  const void* d = (*env)->GetDirectBufferAddress(env, data);
  const uint64_t* o = (const uint64_t*)(*env)->GetDirectBufferAddress(env, offsets);
  const uint64_t* mo = matchOffsets ? (const uint64_t*)(*env)->GetDirectBufferAddress(env, matchOffsets) : NULL;
  jint* c = (*env)->GetPrimitiveArrayCritical(env, counts, NULL);
  jint* s = mo ? (*env)->GetPrimitiveArrayCritical(env, starts, NULL) : NULL;
  jint* e = mo ? (*env)->GetPrimitiveArrayCritical(env, ends, NULL) : NULL;
  int rc = ndl_find_all_batch((ndl_pattern*)(intptr_t)h, d, o, (uint64_t)n, charWidth, (uint32_t*)c, mo, (int32_t*)s, (int32_t*)e,
                              NDL_MEM_HOST, NULL);
  if (e) (*env)->ReleasePrimitiveArrayCritical(env, ends, e, 0);
  if (s) (*env)->ReleasePrimitiveArrayCritical(env, starts, s, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, counts, c, 0);
  if (rc != NDL_OK) throw_for(env, rc);
}

JNIEXPORT jintArray JNICALL Java_com_justinblank_strings_gpu_NeedleNative_matchOne(JNIEnv* env, jclass k, jlong h, jint mode, jstring s, jint from) {
  (void)k;
  jsize n = (*env)->GetStringLength(env, s);
  const jchar* chars = (*env)->GetStringCritical(env, s, NULL);
  uint64_t offsets[2] = {0, (uint64_t)n};
  int32_t f = from, st = -1, en = -1;
  uint8_t m = 0;
  int rc = ndl_match_batch((ndl_pattern*)(intptr_t)h, mode, chars, offsets, 1, 2, &f, &m, &st, &en, NDL_MEM_HOST, NULL);
  (*env)->ReleaseStringCritical(env, s, chars);
  if (rc != NDL_OK) { throw_for(env, rc); return NULL; }
  jint vals[3] = {m, st, en};
  jintArray out = (*env)->NewIntArray(env, 3);
  if (out) (*env)->SetIntArrayRegion(env, out, 0, 3, vals);
  return out;
}

/* findAllStrings: GetStringRegion into one malloc'd UTF-16 buffer + offsets, then one ndl_match_batch. */
JNIEXPORT jobject JNICALL Java_com_justinblank_strings_gpu_NeedleNative_findAllStrings(JNIEnv* env, jclass k, jlong h, jobjectArray hay) {
  (void)k;
  jsize n = (*env)->GetArrayLength(env, hay);
  uint64_t* offsets = (uint64_t*)malloc(((size_t)n + 1) * sizeof(uint64_t));
  uint64_t total = 0;
  offsets[0] = 0;
  for (jsize i = 0; i < n; i++) {
    jstring s = (jstring)(*env)->GetObjectArrayElement(env, hay, i);
    total += (uint64_t)(*env)->GetStringLength(env, s);
    offsets[i + 1] = total;
    (*env)->DeleteLocalRef(env, s);
  }
  jchar* data = (jchar*)malloc((size_t)(total ? total : 1) * sizeof(jchar));
  for (jsize i = 0; i < n; i++) {
    jstring s = (jstring)(*env)->GetObjectArrayElement(env, hay, i);
    (*env)->GetStringRegion(env, s, 0, (jsize)(offsets[i + 1] - offsets[i]), data + offsets[i]);
    (*env)->DeleteLocalRef(env, s);
  }
  jclass rc_cls = (*env)->FindClass(env, "com/justinblank/strings/gpu/GpuPattern$BatchResult");
  jobject res = (*env)->NewObject(env, rc_cls, (*env)->GetMethodID(env, rc_cls, "<init>", "(I)V"), n);
  jbyteArray matched = (jbyteArray)(*env)->GetObjectField(env, res, (*env)->GetFieldID(env, rc_cls, "matched", "[B"));
  jintArray start = (jintArray)(*env)->GetObjectField(env, res, (*env)->GetFieldID(env, rc_cls, "start", "[I"));
  jintArray end = (jintArray)(*env)->GetObjectField(env, res, (*env)->GetFieldID(env, rc_cls, "end", "[I"));
  jbyte* m = (*env)->GetPrimitiveArrayCritical(env, matched, NULL);
  jint* s = (*env)->GetPrimitiveArrayCritical(env, start, NULL);
  jint* e = (*env)->GetPrimitiveArrayCritical(env, end, NULL);
  int rc = ndl_match_batch((ndl_pattern*)(intptr_t)h, NDL_MODE_FIND, data, offsets, (uint64_t)n, 2, NULL, (uint8_t*)m, (int32_t*)s,
                           (int32_t*)e, NDL_MEM_HOST, NULL);
  (*env)->ReleasePrimitiveArrayCritical(env, end, e, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, start, s, 0);
  (*env)->ReleasePrimitiveArrayCritical(env, matched, m, 0);
  free(data);
  free(offsets);
  if (rc != NDL_OK) { throw_for(env, rc); return NULL; }
  return res;
}
