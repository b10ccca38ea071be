/*
 * Minimal stand-in for <jni.h>, used ONLY to syntax-check gatk_b200/csrc/gpuphmm_jni.cpp in the build container,
 * which has no JDK (tests/test_abi_cpu.py::test_jni_shim_syntax).  It declares just the JNI names the shim uses,
 * with the signatures of the JNI specification.  Never shipped, never linked.
 */
#ifndef JNI_STUB_H
#define JNI_STUB_H
#include <cstdint>
#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL
typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef uint8_t jboolean;
#define JNI_FALSE 0
#define JNI_TRUE 1
typedef double jdouble;
typedef jint jsize;
class _jobject {};
class _jclass : public _jobject {};
class _jarray : public _jobject {};
class _jobjectArray : public _jarray {};
class _jbyteArray : public _jarray {};
class _jintArray : public _jarray {};
class _jlongArray : public _jarray {};
class _jdoubleArray : public _jarray {};
typedef _jobject *jobject;
typedef _jclass *jclass;
typedef _jarray *jarray;
typedef _jobjectArray *jobjectArray;
typedef _jbyteArray *jbyteArray;
typedef _jintArray *jintArray;
typedef _jlongArray *jlongArray;
typedef _jdoubleArray *jdoubleArray;
struct _jfieldID;
typedef _jfieldID *jfieldID;
struct JNIEnv {
    jclass FindClass(const char *);
    jint ThrowNew(jclass, const char *);
    jboolean ExceptionCheck();
    jint EnsureLocalCapacity(jint);
    jfieldID GetFieldID(jclass, const char *, const char *);
    jobject GetObjectField(jobject, jfieldID);
    jsize GetArrayLength(jarray);
    jobject GetObjectArrayElement(jobjectArray, jsize);
    void DeleteLocalRef(jobject);
    void GetByteArrayRegion(jbyteArray, jsize, jsize, jbyte *);
    void GetIntArrayRegion(jintArray, jsize, jsize, jint *);
    void SetIntArrayRegion(jintArray, jsize, jsize, const jint *);
    void GetDoubleArrayRegion(jdoubleArray, jsize, jsize, jdouble *);
    void SetDoubleArrayRegion(jdoubleArray, jsize, jsize, const jdouble *);
    void SetByteArrayRegion(jbyteArray, jsize, jsize, const jbyte *);
    void SetLongArrayRegion(jlongArray, jsize, jsize, const jlong *);
    jlongArray NewLongArray(jsize);
    jdoubleArray NewDoubleArray(jsize);
};
#endif
