/* shim: nothing to configure */
