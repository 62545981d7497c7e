/*
 * COMPILE-ONLY stand-in for <jni.h>.  The build image of this repository has no JDK, so tests/test_jni_compiles.py
 * compiles java/src/main/native/needle_jni.c against this file (gcc -Wall -Wextra -Werror) to keep the shim from rotting:
 * types, macros and the JNIEnv functions the shim uses, with the signatures of the JNI specification.  The slots of
 * JNINativeInterface_ are NOT in the specification's order - nothing built against this header may be loaded by a JVM.
 * A real build uses $JAVA_HOME/include/jni.h.
 */
#ifndef NEEDLE_STUB_JNI_H
#define NEEDLE_STUB_JNI_H
#include <stdarg.h>
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
#define JNI_ABORT 2
#define JNI_TRUE 1
#define JNI_FALSE 0

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
typedef uint16_t jchar;
typedef jint jsize;
struct _jobject;
typedef struct _jobject* jobject;
typedef jobject jclass;
typedef jobject jstring;
typedef jobject jthrowable;
typedef jobject jarray;
typedef jarray jbyteArray;
typedef jarray jintArray;
typedef jarray jlongArray;
typedef jarray jobjectArray;
struct _jmethodID;
typedef struct _jmethodID* jmethodID;
struct _jfieldID;
typedef struct _jfieldID* jfieldID;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv*, const char*);
  jint (*ThrowNew)(JNIEnv*, jclass, const char*);
  jint (*Throw)(JNIEnv*, jthrowable);
  jboolean (*ExceptionCheck)(JNIEnv*);
  jthrowable (*ExceptionOccurred)(JNIEnv*);
  void (*ExceptionClear)(JNIEnv*);
  void (*DeleteLocalRef)(JNIEnv*, jobject);
  jobject (*NewObject)(JNIEnv*, jclass, jmethodID, ...);
  jmethodID (*GetMethodID)(JNIEnv*, jclass, const char*, const char*);
  jfieldID (*GetFieldID)(JNIEnv*, jclass, const char*, const char*);
  jobject (*GetObjectField)(JNIEnv*, jobject, jfieldID);
  jstring (*NewStringUTF)(JNIEnv*, const char*);
  jsize (*GetStringLength)(JNIEnv*, jstring);
  void (*GetStringRegion)(JNIEnv*, jstring, jsize, jsize, jchar*);
  jsize (*GetArrayLength)(JNIEnv*, jarray);
  jobject (*GetObjectArrayElement)(JNIEnv*, jobjectArray, jsize);
  jbyteArray (*NewByteArray)(JNIEnv*, jsize);
  jintArray (*NewIntArray)(JNIEnv*, jsize);
  jlongArray (*NewLongArray)(JNIEnv*, jsize);
  void (*SetLongArrayRegion)(JNIEnv*, jlongArray, jsize, jsize, const jlong*);
  void (*GetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, jbyte*);
  void (*SetByteArrayRegion)(JNIEnv*, jbyteArray, jsize, jsize, const jbyte*);
  void (*GetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, jint*);
  void (*SetIntArrayRegion)(JNIEnv*, jintArray, jsize, jsize, const jint*);
  jobject (*NewDirectByteBuffer)(JNIEnv*, void*, jlong);
  void* (*GetDirectBufferAddress)(JNIEnv*, jobject);
  jlong (*GetDirectBufferCapacity)(JNIEnv*, jobject);
};
#endif
